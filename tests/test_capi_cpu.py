"""CPU-side checks of the product library: it loads, exports every symbol the header declares,
and its host loaders (dict.txt in both modes, encoding/gob map[string]int, prob_emit.json, the
math.Log restatement) agree with the oracle.  No compute entry point is called here (no GPU)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from jieba_go_b200 import _capi, synth
from oracle import c_oracle as co
from oracle import py_oracle as po

import kat_vectors as kv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "jieba_b200.h")).read()
    declared = set(re.findall(r"\b(jb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = _capi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "library does not export %s" % name
    assert declared == set(_capi.SYMBOLS), "ctypes table and header disagree: %s" % (declared ^ set(_capi.SYMBOLS))
    assert L.jb_version() == 100


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "jieba_go_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "import oracle" not in src and "from oracle" not in src and "jieba_oracle" not in src, f


def _load_text(data: bytes, mode: int):
    L = _capi.lib()
    db = C.c_void_p()
    buf = (C.c_char * max(1, len(data))).from_buffer_copy(data or b"\0")
    rc = L.jb_dict_load_text(C.cast(buf, C.c_void_p), len(data), mode, C.byref(db))
    return rc, db


def _dump(db):
    L = _capi.lib()
    dd = _capi.DictDesc()
    L.jb_dict_buf_desc(db, C.byref(dd))
    n = dd.n
    off = np.ctypeslib.as_array(C.cast(dd.key_off, C.POINTER(C.c_uint32)), shape=(n + 1,))
    freq = np.ctypeslib.as_array(C.cast(dd.freq, C.POINTER(C.c_int64)), shape=(max(n, 1),))
    blob = C.string_at(dd.keys, int(off[n])) if n else b""
    return {blob[off[i]:off[i + 1]]: int(freq[i]) for i in range(n)}, dd.size


@pytest.mark.parametrize("mode", [0, 1])
def test_dict_text_loader_matches_oracle(small_synth, mode):
    sd, _ = small_synth
    rc, db = _load_text(sd.dict_txt(), mode)
    assert rc == 0
    got, size = _dump(db)
    lines = sd.lines()
    want = po.PrefixDictionary.from_lines_prefix_mode(lines) if mode else po.PrefixDictionary.from_lines_file_mode(lines)
    assert got == want.term_freq and size == want.size
    _capi.lib().jb_dict_buf_free(db)


def test_build_prefix_dict_vector():
    # TestBuildPrefixDict, tokenizer_test.go:431-465
    rc, db = _load_text("\n".join(kv.BUILD_PREFIX_DICT_INPUT).encode(), 1)
    assert rc == 0
    got, size = _dump(db)
    assert {k.decode(): v for k, v in got.items()} == kv.BUILD_PREFIX_DICT_WANT
    L = _capi.lib()
    # TestAddWord, tokenizer_test.go:475-497 (addTerm)
    for term, f in {"左和右": 20, "上和下": 80}.items():
        assert L.jb_dict_add_term(db, term.encode(), len(term.encode()), f) == 0
    v = C.c_int64()
    assert L.jb_dict_buf_lookup(db, "左和右".encode(), 9, C.byref(v)) == 1 and v.value == 20
    assert L.jb_dict_buf_lookup(db, "左和".encode(), 6, C.byref(v)) == 0
    assert _dump(db)[1] == size + 100
    L.jb_dict_buf_free(db)


def test_dict_text_errors():
    rc, _ = _load_text(b"word-without-count\n", 0)   # parts[1] out of range: the reference panics (T:414)
    assert rc == -3 and b"missing frequency" in _capi.lib().jb_last_error()
    rc, _ = _load_text("今天 x1 n\n".encode(), 0)       # strconv.Atoi error: log.Fatal (T:415-417)
    assert rc == -3
    L = _capi.lib()
    db = C.c_void_p()
    assert L.jb_dict_load_file(b"/nonexistent/dict.txt", 0, C.byref(db)) == -2


# ---- encoding/gob -------------------------------------------------------------------------------
def _gob_uint(u):
    if u < 128:
        return bytes([u])
    b = u.to_bytes((u.bit_length() + 7) // 8, "big")
    return bytes([256 - len(b)]) + b


def _gob_int(i):
    return _gob_uint((~i << 1) | 1 if i < 0 else i << 1)


def gob_encode_map_string_int(m, type_id=65):
    """gob.NewEncoder(f).Encode(map[string]int) per the encoding/gob wire spec."""
    # message 1: wireType{MapT: &mapType{CommonType{Name:"", Id:type_id}, Key: 6 (string), Elem: 2 (int)}}
    common = b"\x02" + _gob_int(type_id) + b"\x00"          # CommonType: field 1 (Id) (Name empty omitted -> delta 2)
    mapt = b"\x01" + common + b"\x01" + _gob_int(6) + b"\x01" + _gob_int(2) + b"\x00"
    wire = b"\x04" + mapt + b"\x00"                           # wireType field 3 (MapT): delta 4 from -1
    msg1 = _gob_int(-type_id) + wire
    body = _gob_uint(len(m))
    for k, v in m.items():
        kb = k if isinstance(k, bytes) else k.encode()
        body += _gob_uint(len(kb)) + kb + _gob_int(v)
    msg2 = _gob_int(type_id) + b"\x00" + body
    return _gob_uint(len(msg1)) + msg1 + _gob_uint(len(msg2)) + msg2


def test_gob_loader_roundtrip(small_synth):
    sd, _ = small_synth
    want = po.PrefixDictionary.from_lines_prefix_mode(sd.lines()).term_freq
    want = dict(want)
    want[b"big"] = 2 ** 40 + 123
    data = gob_encode_map_string_int(want)
    L = _capi.lib()
    db = C.c_void_p()
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    assert L.jb_dict_load_gob(C.cast(buf, C.c_void_p), len(data), C.byref(db)) == 0, L.jb_last_error()
    got, size = _dump(db)
    assert got == want and size == 0     # size is not in the gob: the reference hard-codes it (T:454)
    assert po.read_gob_map_string_int(data) == want   # the oracle's independent reader (tests/test_real_data.py uses it)
    L.jb_dict_buf_set_size(db, 60_101_967)
    assert _dump(db)[1] == 60_101_967
    L.jb_dict_buf_free(db)
    bad = data[: len(data) // 2]
    buf = (C.c_char * len(bad)).from_buffer_copy(bad)
    assert L.jb_dict_load_gob(C.cast(buf, C.c_void_p), len(bad), C.byref(db)) == -3


# ---- prob_emit.json -----------------------------------------------------------------------------
def test_emit_json_loader(small_synth):
    from jieba_go_b200.tokenizer import _load_emit_bytes
    sd, emit = small_synth
    for ensure_ascii in (False, True):   # raw UTF-8 keys and \\uXXXX escapes
        obj = {s: {chr(c): v for c, v in tab.items()} for s, tab in emit.items()}
        obj["S"]["\U00020000"] = -7.25    # surrogate-pair escape when ensure_ascii
        obj["B"]["ab"] = -1.0             # multi-rune keys are never queried (T:689,708): skipped
        data = json.dumps(obj, ensure_ascii=ensure_ascii).encode()
        st, ru, lp = _load_emit_bytes(data)
        got = {}
        for s, r, v in zip(st.tolist(), ru.tolist(), lp.tolist()):
            got[("BMES"[s], r)] = v
        want = {(s, c): v for s, tab in emit.items() for c, v in tab.items()}
        want[("S", 0x20000)] = -7.25
        assert got == want   # identical float64 bits (strtod is correctly rounded like Go's ParseFloat)
    # TestLoadHMM constants (tokenizer_test.go:291-300) survive the parser bit for bit
    data = json.dumps({s: {"一": v} for s, v in kv.LOAD_HMM.items()}).encode()
    st, ru, lp = _load_emit_bytes(data)
    assert {"BMES"[s]: v for s, v in zip(st.tolist(), lp.tolist())} == kv.LOAD_HMM
    with pytest.raises(_capi.JiebaB200Error):
        _load_emit_bytes(b'{"B": {"x": }')


def test_go_log_matches_oracle():
    L = _capi.lib()
    rng = np.random.default_rng(2)
    xs = np.concatenate([np.arange(1, 30000), rng.integers(1, 2 ** 40, 30000)]).astype(np.float64)
    for x in xs.tolist():
        assert L.jb_go_log(x) == co.go_log(x)
    assert L.jb_go_log(0.0) == -np.inf


def test_hmm_defaults_match_reference_literals():
    L = _capi.lib()
    hd = _capi.HmmDesc()
    L.jb_hmm_defaults(C.byref(hd))
    ref = po.HiddenMarkovModel({})
    assert [hd.start[i] for i in range(4)] == [ref.start_p[s] for s in "BMES"]
    for a, p in enumerate("BMES"):
        for b, n in enumerate("BMES"):
            want = ref.trans_p.get(p, {}).get(n, 0.0)
            assert hd.trans[a][b] == want


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from jieba_go_b200.tokenizer import Tokenizer
    with pytest.raises(_capi.JiebaB200Error) as ei:
        Tokenizer.from_dict_text("甲 1\n".encode(), 1, {"B": {}, "M": {}, "E": {}, "S": {}})
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


# ---- negative counts: rejected everywhere (include/jieba_b200.h, Limits) ---------------------------
def test_negative_counts_are_rejected():
    L = _capi.lib()
    for mode in (0, 1):
        rc, _ = _load_text("甲 5\n乙 -3 n\n".encode(), mode)
        assert rc == -3 and b"negative" in L.jb_last_error()
        with pytest.raises(ValueError):
            co.Dict.from_lines("甲 5\n乙 -3 n\n", mode)
    with pytest.raises(ValueError):
        po.PrefixDictionary.from_lines_prefix_mode(["甲 5", "乙 -3 n"])
    with pytest.raises(ValueError):
        po.PrefixDictionary.from_lines_file_mode(["甲 5", "乙 -3 n"])
    data = gob_encode_map_string_int({b"a": 1, b"neg": -7})
    db = C.c_void_p()
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    assert L.jb_dict_load_gob(C.cast(buf, C.c_void_p), len(data), C.byref(db)) == -3
    rc, db = _load_text("甲 5\n".encode(), 1)
    assert rc == 0
    assert L.jb_dict_add_term(db, "乙".encode(), 3, -1) == -1
    L.jb_dict_buf_free(db)


# ---- suggestFreq (T:589-614): the library's arithmetic against the oracle's restatement ---------------
def test_suggest_freq_matches_oracle(small_synth):
    sd, _ = small_synth
    L = _capi.lib()
    rc, db = _load_text(sd.dict_txt(), 1)
    assert rc == 0
    pd = po.PrefixDictionary.from_lines_prefix_mode(sd.lines())
    rng = np.random.default_rng(21)
    words = [w for w in sd.words]
    for trial in range(300):
        k = int(rng.integers(1, 6))
        pieces = [words[int(i)] for i in rng.integers(0, len(words), k)]
        if trial % 7 == 0:
            pieces.append("龥龥".encode())          # a piece that is not a key: counts as 1
        term = b"".join(pieces) if trial % 3 else words[int(rng.integers(0, len(words)))]
        off = np.zeros(len(pieces) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(x) for x in pieces])
        out = C.c_int64()
        assert L.jb_dict_suggest_freq(db, term, len(term), b"".join(pieces), off.ctypes.data, len(pieces), C.byref(out)) == 0
        assert out.value == pd.suggest_freq(term, pieces), (term, pieces)
    # an empty dictionary: size < 1 counts as 1 (T:590-593)
    rc, e = _load_text(b"", 1)
    out = C.c_int64()
    off = np.array([0, 3], dtype=np.uint64)
    assert L.jb_dict_suggest_freq(e, "甲".encode(), 3, "甲".encode(), off.ctypes.data, 1, C.byref(out)) == 0
    assert out.value == po.PrefixDictionary().suggest_freq("甲".encode(), ["甲".encode()]) == 2
    L.jb_dict_buf_free(db)
    L.jb_dict_buf_free(e)


def test_host_pool_limit_without_device():
    L = _capi.lib()
    assert L.jb_host_pool_limit(0) == 0       # nothing pooled yet; releases everything
    assert L.jb_host_pool_limit(4 << 30) == 0


# ---- the real-data harness, exercised on synthetic stand-ins ------------------------------------------
def test_real_data_harness_mechanics(tmp_path, monkeypatch, small_synth):
    """tests/test_real_data.py only runs where the reference's LFS files exist.  Here its locator and fixtures run on
    synthetic files whose digests are patched in, so that the harness itself is known to work."""
    import hashlib
    import realdata
    import test_real_data as trd
    sd, emit = small_synth
    pd = po.PrefixDictionary.from_lines_prefix_mode(sd.lines())
    files = {"dict.txt": sd.dict_txt(), "prefix_dictionary.gob": gob_encode_map_string_int(pd.term_freq),
             "prob_emit.json": synth.emit_json(emit)}
    for name, data in files.items():
        (tmp_path / name).write_bytes(data)
    monkeypatch.setattr(kv, "REAL_SHA256", {n: hashlib.sha256(d).hexdigest() for n, d in files.items()})
    monkeypatch.setenv("JIEBA_DATA_DIR", str(tmp_path))
    monkeypatch.setattr(realdata, "_cache", None)
    paths, why = realdata.locate()
    assert why is None and set(paths) == set(files)
    paths, emit2, gob, ptk, ctk = trd.build_oracles(paths)
    assert gob == pd.term_freq
    text = (sd.words[3] + sd.words[10] + "，龥".encode() + sd.words[4]).decode()
    for hmm in (False, True):
        assert ptk.cut_strings(text, hmm) == ctk.cut_strings(text, hmm)
    runes = [ord(c) for c in sd.words[3].decode() + sd.words[10].decode()]
    assert ptk.pd.build_dag(runes) == ctk.build_dag("".join(map(chr, runes)))
    assert ptk.hmm.viterbi(runes) == ctk.hmm.viterbi("".join(map(chr, runes)))
    monkeypatch.setattr(realdata, "_cache", None)


# ---- cached table image (jb_tokenizer_create_cached) ------------------------------------------------
def test_sha256_matches_hashlib():
    import hashlib
    L = _capi.lib()
    rng = np.random.default_rng(5)
    for n in [0, 1, 3, 55, 56, 57, 63, 64, 65, 119, 120, 1000, 100_003]:
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        out = C.create_string_buffer(32)
        L.jb_debug_sha256(data, n, out)
        assert out.raw == hashlib.sha256(data).digest(), n


def test_table_image_cache_lifecycle_without_device(tmp_path, small_synth):
    """No GPU here: every call ends with JB_ECUDA at the upload, but the host half (parse, build, write / read back the
    image, key check) has run by then, and from_cache tells which way it went."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present (tests/test_gpu_parity.py covers the cached constructor there)")
    sd, emit = small_synth
    L = _capi.lib()
    dp, ep, ip = tmp_path / "dict.txt", tmp_path / "prob_emit.json", tmp_path / "tables.img"
    dp.write_bytes(sd.dict_txt())
    ep.write_bytes(synth.emit_json(emit))

    def create(kind=1):
        h = C.c_void_p()
        used = C.c_int(-1)
        rc = L.jb_tokenizer_create_cached(str(dp).encode(), kind, 0, str(ep).encode(), None, str(ip).encode(), C.byref(used), C.byref(h))
        assert rc == -4, L.jb_last_error()   # JB_ECUDA: no device
        return used.value

    assert create() == 0 and ip.exists()          # built from the files, image written
    size = ip.stat().st_size
    assert size > 65536 * 16
    assert create() == 1                          # read back
    assert create(kind=0) == 0                    # another loader mode: another key, rebuilt and rewritten
    assert create(kind=0) == 1
    assert create() == 0                          # ... and back
    raw = bytearray(ip.read_bytes())
    raw[len(raw) // 2] ^= 0x40                    # damage the payload
    ip.write_bytes(bytes(raw))
    assert create() == 0
    assert create() == 1
    ip.write_bytes(ip.read_bytes()[: size // 3])  # truncate
    assert create() == 0
    dp.write_bytes(sd.dict_txt() + "\u9f98\u9f98 7 n\n".encode())   # the dictionary changed: stale image
    assert create() == 0
    assert create() == 1
