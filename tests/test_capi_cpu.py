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
    want[b"neg"] = -7
    want[b"big"] = 2 ** 40 + 123
    data = gob_encode_map_string_int(want)
    L = _capi.lib()
    db = C.c_void_p()
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    assert L.jb_dict_load_gob(C.cast(buf, C.c_void_p), len(data), C.byref(db)) == 0, L.jb_last_error()
    got, size = _dump(db)
    assert got == want and size == 0     # size is not in the gob: the reference hard-codes it (T:454)
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
