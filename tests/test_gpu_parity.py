"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same
inputs.  Bit-exact bar: token (start,end) lists and doc_tok_off arrays must be identical; route
values (float64) must have identical bits."""
import numpy as np
import pytest

from jieba_go_b200 import synth

import kat_vectors as kv
from helpers import c_oracle_tokenizer, emit_arrays, fuzz_docs, pack_docs  # noqa: F401

pytestmark = pytest.mark.gpu


PATHS = ["stream", "general"]  # default streaming fast path (k_scan/k_route/k_emit) / general kernels only


def _gpu_tokenizer(sd_or_lines, emit, mode=1, path="stream", **kw):
    from jieba_go_b200.tokenizer import Tokenizer
    if isinstance(sd_or_lines, (list, tuple)):
        data = "\n".join(sd_or_lines).encode() + b"\n"
    else:
        data = sd_or_lines.dict_txt()
    tk = Tokenizer.from_dict_text(data, mode, emit, **kw)
    if path == "general":
        tk.set_general_only(True)
    return tk


@pytest.fixture(scope="module", params=PATHS)
def kat_tk(request, kat_lines, kat_emit):
    return _gpu_tokenizer(kat_lines, kat_emit, path=request.param)


@pytest.fixture(scope="module")
def synth_pair(small_synth):
    sd, emit = small_synth
    return sd, emit, _gpu_tokenizer(sd, emit), c_oracle_tokenizer(sd, emit)


@pytest.fixture(scope="module", params=PATHS)
def medium_pair(request):
    sd = synth.make_dictionary(n_words=60000, seed=synth.SEED_BASE + 31, total_freq=1.5e7, max_len=16)
    emit = synth.make_emit(sd, seed=synth.SEED_BASE + 32)
    return sd, emit, _gpu_tokenizer(sd, emit, path=request.param), c_oracle_tokenizer(sd, emit)


def _assert_same(gpu, ora, text=None, off=None):
    gs, ge, gd = gpu
    os_, oe, _, od = ora
    assert np.array_equal(gd, od), "doc_tok_off differs"
    if not (np.array_equal(gs, os_) and np.array_equal(ge, oe)):
        n = min(len(gs), len(os_))
        bad = np.nonzero((gs[:n] != os_[:n]) | (ge[:n] != oe[:n]))[0]
        i = int(bad[0]) if len(bad) else n
        d = int(np.searchsorted(od, i, side="right") - 1)
        ctx = ""
        if text is not None:
            doc = bytes(text[int(off[d]):int(off[d + 1])])
            lo = max(0, int(os_[max(i - 2, int(od[d]))]) if i < len(os_) else 0)
            ctx = repr(doc[lo:lo + 60].decode("utf-8", "replace"))
        raise AssertionError("token %d (doc %d) differs: gpu=%s oracle=%s near %s; counts gpu=%d oracle=%d" % (
            i, d, (gs[i:i + 4].tolist(), ge[i:i + 4].tolist()), (os_[i:i + 4].tolist(), oe[i:i + 4].tolist()), ctx, len(gs), len(os_)))


# ---- reference vectors + micro-KATs through the GPU path ---------------------------------------
@pytest.mark.parametrize("text,off,on", kv.KATS)
def test_kats(kat_tk, text, off, on):
    assert kat_tk.cut(text, False) == off
    assert kat_tk.cut(text, True) == on


def test_kat5_selector_is_not_argmax(kat_emit):
    tk = _gpu_tokenizer(kv.KAT5_LINES, kat_emit)
    assert tk.cut(kv.KAT5[0], False) == kv.KAT5[1]


@pytest.mark.parametrize("text,want", kv.CUT_NON_ZH)  # TestCutNonZh, tokenizer_test.go:373-376
def test_cut_non_zh(kat_tk, text, want):
    assert kat_tk.cut(text, False) == want


@pytest.mark.parametrize("text,want", kv.SPLIT_TEXT)  # TestSplitText: block structure seen through Cut
def test_split_text_blocks(kat_tk, kat_lines, kat_emit, text, want):
    ora = c_oracle_tokenizer_lines(kat_lines, kat_emit)
    assert kat_tk.cut_offsets(text, True) == ora.cut(text, True)
    assert kat_tk.cut_offsets(text, False) == ora.cut(text, False)


def c_oracle_tokenizer_lines(lines, emit):
    from oracle import c_oracle as co
    return co.Tokenizer(co.Dict.from_lines(lines, 1), co.Hmm(emit))


def test_cut_parallel_contract(kat_tk, kat_lines, kat_emit):
    from oracle import py_oracle as po
    ora = po.Tokenizer(po.PrefixDictionary.from_lines_prefix_mode(kat_lines), po.HiddenMarkovModel(kat_emit))
    t = "乙丙，a1 乙丙甲甲甲x 乙\t丙丁!"
    want = ora.cut_strings(t, True)
    assert want == ["乙丙", "，", "a1", "乙丙", "甲甲", "甲", "x", "乙", "丙丁"]
    assert kat_tk.cut_parallel(t, True, 6, True) == want                  # ordered=true == Cut (T:110-125)
    assert sorted(kat_tk.cut_parallel(t, True, 6, False)) == sorted(want)  # any block order (T:126-133)


def test_invalid_utf8(kat_tk, kat_lines, kat_emit):
    b = b"a\xff\xe4\xb8 \xe4\xb9\x99\x80z"
    want = [(0, 1, False), (1, 2, True), (2, 3, True), (3, 4, True), (5, 8, False), (8, 9, True), (9, 10, False)]
    assert kat_tk.cut_offsets(b, False) == want
    assert kat_tk.cut(b, False) == ["a", "�", "�", "�", "乙", "�", "z"]


# ---- dictionary table ---------------------------------------------------------------------------
def test_device_dictionary_matches_term_freq(synth_pair):
    sd, emit, tk, ora = synth_pair
    from oracle import c_oracle as co
    keys, koff, freq = ora.pd.export()
    log_total = co.go_log(float(ora.pd.size))
    rng = np.random.default_rng(3)
    idx = rng.choice(len(freq), size=min(400, len(freq)), replace=False)
    n_han = 0
    for i in idx.tolist():
        k = keys[koff[i]:koff[i + 1]].tobytes()
        s = k.decode()
        if not all(0x4E00 <= ord(ch) <= 0x9FFF for ch in s):
            continue
        n_han += 1
        kind, w = tk.debug_lookup(k)
        if freq[i] > 0:
            assert kind == 2 and w == co.go_log(float(freq[i])) - log_total, s
        else:
            assert kind == 1, s
    assert n_han > 100
    assert tk.debug_lookup("龥龥龥")[0] == 0


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("runes", [1, 2, 7, 300, 1024, 1025, 3000])
def test_route_values_bit_exact(small_synth, path, runes):
    """float64 route values R[i] = maxIndexProba(dagProba[i]) (T:502-548, 565-578), bit for bit, from the kernel that
    cuts the block on each path: k_route (`stream`: the kernel behind every benchmark number) and k_route_dp (`general`)."""
    sd, emit = small_synth
    tk = _gpu_tokenizer(sd, emit, 1, path=path)
    ora = c_oracle_tokenizer(sd, emit, 1)
    text, _ = synth.make_corpus(sd, "long", 40_000, synth.SEED_BASE + 40)
    han = text.numpy()[:3 * runes].tobytes()  # Han only
    ge, gp = tk.debug_route(han)
    oe, op = ora.route(han)
    assert len(ge) == runes
    assert np.array_equal(ge, oe)
    assert np.array_equal(gp.view(np.uint64), op.view(np.uint64))


# ---- randomised parity ---------------------------------------------------------------------------
@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("hmm", [False, True])
def test_fuzz_docs(small_synth, mode, hmm, path):
    sd, emit = small_synth
    tk = _gpu_tokenizer(sd, emit, mode, path=path)
    ora = c_oracle_tokenizer(sd, emit, mode)
    rng = np.random.default_rng(100 + mode)
    docs = fuzz_docs(sd, rng, n_docs=400, max_len=150) + [b"", b"", "甲".encode(), b"\xe4", b"x"]
    text, off = pack_docs(docs)
    _assert_same(tk.cut_batch(text, off, hmm), ora.cut_batch(text, off, hmm, 2), text, off)


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("hmm", [False, True])
def test_fuzz_docs_streaming_path(small_synth, mode, hmm):
    """No 4-byte Han rune anywhere, so the batch really stays on k_scan/k_route/k_emit."""
    sd, emit = small_synth
    tk = _gpu_tokenizer(sd, emit, mode)
    ora = c_oracle_tokenizer(sd, emit, mode)
    rng = np.random.default_rng(300 + mode)
    docs = fuzz_docs(sd, rng, n_docs=1500, max_len=150, supp_han=False) + [b"", b"", "甲".encode(), b"\xe4", b"x"]
    text, off = pack_docs(docs)
    _assert_same(tk.cut_batch(text, off, hmm), ora.cut_batch(text, off, hmm, 2), text, off)


@pytest.mark.parametrize("kind", ["freq", "oov", "long"])
@pytest.mark.parametrize("hmm", [False, True])
def test_corpora(medium_pair, kind, hmm):
    sd, emit, tk, ora = medium_pair
    text, doc_off = synth.make_corpus(sd, kind, 3_000_000, synth.SEED_BASE + 50)
    t = text.numpy()
    off = doc_off.numpy().astype(np.uint64)
    _assert_same(tk.cut_batch(t, off, hmm), ora.cut_batch(t, off, hmm, 8), t, off)


@pytest.mark.parametrize("path", PATHS)
def test_unicode_13_vs_15(small_synth, path):
    sd, emit = small_synth
    t = "鿽鿾鿿𪛞乙".encode()  # U+9FFD..9FFF, U+2A6DE: Han only from Unicode 14/15 on
    for ver in (13, 15):
        tk = _gpu_tokenizer(sd, emit, 1, path=path, unicode_version=ver)
        ora = c_oracle_tokenizer(sd, emit, 1, unicode_version=ver)
        assert tk.cut_offsets(t, True) == ora.cut(t, True)
        assert tk.cut_offsets(t, False) == ora.cut(t, False)


@pytest.mark.parametrize("path", PATHS)
def test_batches_and_document_boundaries(synth_pair, path):
    sd, emit, _, ora = synth_pair
    tk = _gpu_tokenizer(sd, emit, 1, path=path, max_batch_bytes=20_000)  # forces many device batches
    text, doc_off = synth.make_corpus(sd, "oov", 300_000, synth.SEED_BASE + 60)
    t = text.numpy()
    # re-cut the same bytes into documents at arbitrary byte positions (splits runes and words)
    rng = np.random.default_rng(5)
    cuts = np.unique(np.concatenate([[0, t.size], rng.integers(0, t.size, 400)])).astype(np.uint64)
    off = np.concatenate([cuts[:50], cuts[49:50], cuts[49:50], cuts[50:]])  # a few empty documents too
    for hmm in (False, True):
        _assert_same(tk.cut_batch(t, off, hmm), ora.cut_batch(t, off, hmm, 4), t, off)


def test_add_word_rebuilds_tables(kat_lines, kat_emit):
    tk = _gpu_tokenizer(kat_lines, kat_emit)
    assert tk.cut("甲乙", False) == ["甲", "乙"]
    tk.add_word("甲乙", 5000)   # addTerm: termFreq + size, no prefixes (T:580-585)
    assert tk.lookup("甲乙") == 5000 and tk.size == 359 + 5000
    assert tk.cut("甲乙", False) == ["甲乙"]


def test_device_api(medium_pair):
    import torch
    sd, emit, tk, ora = medium_pair
    text, doc_off = synth.make_corpus(sd, "oov", 2_000_000, synth.SEED_BASE + 70)
    t = text.numpy()
    off = doc_off.numpy().astype(np.uint64)
    os_, oe, _, od = ora.cut_batch(t, off, True, 8)
    dt, ddo = text.cuda(), doc_off.cuda()
    cap = len(os_) + 10
    d_start = torch.zeros(cap, dtype=torch.int32, device="cuda")
    d_end = torch.zeros(cap, dtype=torch.int32, device="cuda")
    d_dto = torch.zeros(off.size, dtype=torch.int64, device="cuda")
    d_nt = torch.zeros(2, dtype=torch.int64, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        tk.cut_device(dt, ddo, True, d_start, d_end, d_dto, d_nt, stream=s)
    s.synchronize()
    assert d_nt.tolist() == [len(os_), 0]
    assert np.array_equal(d_start[:len(os_)].cpu().numpy().astype(np.uint32), os_)
    assert np.array_equal(d_end[:len(os_)].cpu().numpy().astype(np.uint32), oe)
    assert np.array_equal(d_dto.cpu().numpy().astype(np.uint64), od)


# ---- routes specific to the streaming fast path -------------------------------------------------
def _mixed_corpus(sd, rng, nbytes):
    """Dictionary words with every kind of interruption the fast path hands over or defers: long
    unpunctuated blocks, Japanese/Korean runs (gated tokens across tile edges), ASCII, 2- and 4-byte
    runes, ill-formed bytes, a 4-byte Han rune far into the text."""
    words = [w for w in sd.words if all(0x4E00 <= ord(c) <= 0x9FA5 for c in w.decode())]
    parts, size = [], 0
    kana = "ステーションかきくけこ번역하다".encode()
    while size < nbytes:
        r = rng.random()
        if r < 0.02:      # long block: 400-3000 runes without a separator
            p = b"".join(words[int(i)] for i in rng.integers(0, len(words), int(rng.integers(150, 1200))))
        elif r < 0.06:    # long non-Han, non-alnum run (crosses tiles, no alnum => dropped)
            p = kana * int(rng.integers(5, 200))
        elif r < 0.10:    # same with one alnum somewhere far away => kept
            p = kana * int(rng.integers(5, 150)) + b"x" + kana * int(rng.integers(0, 150))
        elif r < 0.14:
            p = [" a1b2 ", "é", "€", "\U0001F600", "\n", "\t", "　", "+=", "ＡＢ"][int(rng.integers(0, 9))].encode()
        elif r < 0.15:
            p = [b"\xff", b"\xe4\xb8", b"\x80", b"\xf0\x9f"][int(rng.integers(0, 4))]
        elif r < 0.30:
            p = ["，", "。", "！", "？", "；"][int(rng.integers(0, 5))].encode()
        else:
            p = words[int(rng.integers(0, len(words)))]
        parts.append(p)
        size += len(p)
    return b"".join(parts)


@pytest.mark.parametrize("hmm", [False, True])
def test_fast_path_handover_routes(medium_pair, hmm):
    sd, emit, tk, ora = medium_pair
    rng = np.random.default_rng(77)
    docs = [_mixed_corpus(sd, rng, 150_000) for _ in range(6)]
    text, off = pack_docs(docs)
    _assert_same(tk.cut_batch(text, off, hmm), ora.cut_batch(text, off, hmm, 8), text, off)
    # a 4-byte Han rune sends its block to k_wide; results must not change
    docs2 = docs[:2] + ["甲\U00020000乙".encode() + docs[2]] + docs[3:]
    text, off = pack_docs(docs2)
    _assert_same(tk.cut_batch(text, off, hmm), ora.cut_batch(text, off, hmm, 8), text, off)


@pytest.mark.parametrize("path", PATHS)
def test_long_single_rune_runs(small_synth, path):
    """Viterbi over runs of out-of-vocabulary runes of every length around the 16-rune register window of
    k_emit (shorter runs keep their back-pointers in registers, longer ones in HBM)."""
    sd, emit = small_synth
    tk = _gpu_tokenizer(sd, emit, 1, path=path)
    ora = c_oracle_tokenizer(sd, emit, 1)
    rng = np.random.default_rng(11)
    known = set(w.decode() for w in sd.words if len(w) == 3)
    pool = [chr(c) for c in range(0x4E00, 0x9FA6) if chr(c) not in known]
    docs = []
    for n in list(range(1, 41)) + [63, 64, 65, 100, 257, 1000, 5000]:
        run = "".join(pool[int(i)] for i in rng.integers(0, len(pool), n))
        docs.append(run.encode())
        docs.append((sd.words[3].decode() + run + "，" + run[: n // 2] + sd.words[5].decode() + run).encode())
    text, off = pack_docs(docs)
    for hmm in (True, False):
        _assert_same(tk.cut_batch(text, off, hmm), ora.cut_batch(text, off, hmm, 2), text, off)


@pytest.mark.parametrize("hmm", [False, True])
def test_long_blocks_cut_into_segments(medium_pair, hmm):
    """Han blocks of 512 runes or more are walked segment by segment (k_land / k_chain / k_emit over segments / k_runs):
    block lengths on either side of the threshold and of every segment-count edge, made of dictionary words (tokens of
    2..16 runes that straddle the 256-rune boundaries), of unknown runes (single-rune pieces: Viterbi runs that cross
    the boundaries, also longer than the 24-rune register window) and of both mixed; blocks side by side in one
    document (separated by ASCII only) and several long blocks per batch."""
    sd, emit, tk, ora = medium_pair
    rng = np.random.default_rng(77 + int(hmm))
    known = set(w.decode() for w in sd.words if len(w) == 3)
    pool = [chr(c) for c in range(0x4E00, 0x9FA6) if chr(c) not in known]
    words = [w.decode() for w in sd.words]

    def block(n, p_unknown):
        out, have = [], 0
        while have < n:
            if rng.random() < p_unknown:
                k = int(rng.integers(1, 40)) if rng.random() < 0.1 else 1
                piece = "".join(pool[int(i)] for i in rng.integers(0, len(pool), k))
            else:
                piece = words[int(rng.integers(0, len(words)))]
            piece = piece[: n - have]
            out.append(piece)
            have += len(piece)
        return "".join(out)

    docs = []
    for n in [511, 512, 513, 767, 768, 769, 1023, 1024, 1025, 1279, 1280, 1281, 2048, 4097, 8192 + 255, 8192 + 256, 30000]:
        for pu in (0.0, 0.3, 1.0):
            docs.append(block(n, pu).encode())
    docs.append((block(700, 0.2) + "a" + block(600, 0.5) + " " + block(100, 0.3) + "," + block(513, 0.0)).encode())
    docs.append(("x" * 7 + block(5000, 0.3)).encode())  # the block does not start on a multiple of 3
    text, off = pack_docs(docs)
    _assert_same(tk.cut_batch(text, off, hmm), ora.cut_batch(text, off, hmm, 4), text, off)
    # and through the small-call path (one document below 8 KiB with a long block in it)
    one = block(2000, 0.3).encode()
    t1, o1 = pack_docs([one])
    _assert_same(tk.cut_batch(t1, o1, hmm), ora.cut_batch(t1, o1, hmm, 1), t1, o1)


@pytest.mark.parametrize("path", PATHS)
def test_keys_longer_than_16_runes(path):
    """A 20- and a 30-rune key: the route ring has 32 cells and path entries take 8 bits (k_route<32,8>)."""
    sd = synth.make_dictionary(n_words=3000, seed=synth.SEED_BASE + 91, total_freq=1.0e6, max_len=8)
    emit = synth.make_emit(sd, seed=synth.SEED_BASE + 92)
    lines = [ln.decode() for ln in sd.lines()]
    w20 = "".join(chr(0x4E00 + 7 * i) for i in range(20))
    w30 = "".join(chr(0x5E00 + 11 * i) for i in range(30))
    lines = lines + ["%s 900 n" % w20, "%s 70000 n" % w30]
    tk = _gpu_tokenizer(lines, emit, 1, path=path)
    from oracle import c_oracle as co
    hm = co.Hmm()
    hm.set_emit_arrays(*emit_arrays(emit))
    ora = co.Tokenizer(co.Dict.from_lines(lines, 1), hm)
    rng = np.random.default_rng(12)
    docs = []
    for _ in range(200):
        parts = []
        for _ in range(int(rng.integers(1, 40))):
            r = rng.random()
            if r < 0.1:
                parts.append(w20[: int(rng.integers(1, 21))])
            elif r < 0.2:
                parts.append(w30[int(rng.integers(0, 5)): int(rng.integers(5, 31))])
            elif r < 0.25:
                parts.append("，")
            else:
                parts.append(sd.words[int(rng.integers(0, len(sd.words)))].decode())
        docs.append("".join(parts).encode())
    text, off = pack_docs(docs)
    for hmm in (False, True):
        _assert_same(tk.cut_batch(text, off, hmm), ora.cut_batch(text, off, hmm, 2), text, off)


@pytest.mark.parametrize("hmm", [False, True])
def test_scan_tile_edges(small_synth, hmm):
    """k_scan works on 8128-byte tiles with a 32-byte halo word on either side: put every kind of content on
    the edges (text lengths and document starts within a few bytes of k * 8128)."""
    sd, emit = small_synth
    tk = _gpu_tokenizer(sd, emit, 1)
    ora = c_oracle_tokenizer(sd, emit, 1)
    rng = np.random.default_rng(13)
    words = [w for w in sd.words]
    fillers = ["，", "a1", " ", "é", "€", "ステ", "\n", "9"]
    docs = []
    for tiles in (1, 2, 3):
        for delta in range(-5, 6):
            target = tiles * 8128 + delta
            parts, size = [], 0
            while size < target - 40:
                p = words[int(rng.integers(0, len(words)))] if rng.random() < 0.9 else fillers[int(rng.integers(0, len(fillers)))].encode()
                parts.append(p)
                size += len(p)
            tail = [b"a", "，".encode(), "乙".encode(), b" ", "é".encode()]
            while size < target:
                p = tail[int(rng.integers(0, len(tail)))]
                if size + len(p) > target:
                    p = b"x"
                parts.append(p)
                size += len(p)
            docs.append(b"".join(parts))
    # one document per batch (its end is the end of the text) ...
    for d in docs:
        t = np.frombuffer(d, dtype=np.uint8)
        off = np.array([0, len(d)], dtype=np.uint64)
        _assert_same(tk.cut_batch(t, off, hmm), ora.cut_batch(t, off, hmm, 1), t, off)
    # ... and all of them in one batch (document starts near tile edges)
    text, off = pack_docs(docs)
    _assert_same(tk.cut_batch(text, off, hmm), ora.cut_batch(text, off, hmm, 2), text, off)


@pytest.mark.parametrize("path", PATHS)
def test_four_byte_han_blocks(path):
    """Han blocks with 4-byte runes (CJK extension B): k_route hands them to k_wide.  Dictionary words with such runes,
    runs of them (Viterbi with supplementary-plane emissions), blocks that start / end with one, document boundaries
    next to them, many of them in one batch."""
    sd = synth.make_dictionary(n_words=3000, seed=synth.SEED_BASE + 93, total_freq=1.0e6, max_len=6)
    emit = synth.make_emit(sd, seed=synth.SEED_BASE + 94)
    ext = [chr(0x20000 + 37 * i) for i in range(40)]
    emit = {st: dict(tab) for st, tab in emit.items()}
    for i, c in enumerate(ext[:25]):  # some of them have emissions, in some states
        for j, st in enumerate("BMES"):
            if (i + j) % 3:
                emit[st][ord(c)] = -3.0 - 0.37 * i - 0.11 * j
    lines = [ln.decode() for ln in sd.lines()]
    w = [x.decode() for x in sd.words]
    lines += ["%s 800 n" % ext[0], "%s%s 500 n" % (ext[1], w[3][:1]), "%s%s%s 900 n" % (w[5][:1], ext[2], ext[3]),
              "%s%s 300 n" % (ext[4], ext[5]), "%s%s%s%s 100 n" % (ext[4], ext[5], w[7][:1], ext[6])]
    tk = _gpu_tokenizer(lines, emit, 1, path=path)
    from oracle import c_oracle as co
    hm = co.Hmm()
    hm.set_emit_arrays(*emit_arrays(emit))
    ora = co.Tokenizer(co.Dict.from_lines(lines, 1), hm)
    rng = np.random.default_rng(14)
    docs = []
    for _ in range(400):
        parts = []
        for _ in range(int(rng.integers(1, 30))):
            r = rng.random()
            if r < 0.25:
                parts.append(ext[int(rng.integers(0, len(ext)))])
            elif r < 0.30:
                parts.append("".join(ext[int(i)] for i in rng.integers(0, len(ext), int(rng.integers(2, 45)))))
            elif r < 0.40:
                parts.append(chr(int(rng.integers(0x4E00, 0x9FA6))))
            elif r < 0.47:
                parts.append(["，", "a1", " ", "\n", "é"][int(rng.integers(0, 5))])
            else:
                parts.append(w[int(rng.integers(0, len(w)))])
        docs.append("".join(parts).encode())
    docs += [ext[0].encode(), (ext[1] + w[3][:1]).encode(), (w[2] + ext[7]).encode(), (ext[8] + w[2]).encode()]
    text, off = pack_docs(docs)
    for hmm in (False, True):
        _assert_same(tk.cut_batch(text, off, hmm), ora.cut_batch(text, off, hmm, 2), text, off)
    # the same bytes as ONE document and cut at arbitrary positions (splits the 4-byte runes too)
    cuts = np.unique(np.concatenate([[0, text.size], rng.integers(0, text.size, 300)])).astype(np.uint64)
    for hmm in (False, True):
        _assert_same(tk.cut_batch(text, cuts, hmm), ora.cut_batch(text, cuts, hmm, 2), text, cuts)


@pytest.mark.parametrize("hmm", [False, True])
def test_list_overflow_falls_back_to_general_kernels(small_synth, hmm):
    """A megabyte of punctuation without any alnum: every rune is a gated token whose block leaves its k_scan tile, the
    deferred list overflows, the batch is flagged on the device and the general kernels redo it (same results)."""
    sd, emit = small_synth
    tk = _gpu_tokenizer(sd, emit, 1)
    ora = c_oracle_tokenizer(sd, emit, 1)
    w = [x.decode() for x in sd.words[:50]]
    docs = [("。，！？" * 90_000).encode(),                                   # dropped entirely (T:291-293)
            ("。，！？" * 90_000 + "x" + "；" * 1000).encode(),                # one alnum far away: every rune is a token
            ("".join(w) + "，" * 50_000 + "".join(w[::-1]) + " a1 " + "。" * 40_000).encode()]
    for d in docs:
        t = np.frombuffer(d, dtype=np.uint8)
        off = np.array([0, len(d)], dtype=np.uint64)
        _assert_same(tk.cut_batch(t, off, hmm), ora.cut_batch(t, off, hmm, 4), t, off)
    text, off = pack_docs(docs)
    _assert_same(tk.cut_batch(text, off, hmm), ora.cut_batch(text, off, hmm, 4), text, off)


# ---- the benchmark dictionary (349k words) at a size the oracle still finishes in seconds ---------------------
@pytest.fixture(scope="module")
def bench_pair():
    sd = synth.make_dictionary(n_words=349_000, seed=synth.SEED_BASE)
    emit = synth.make_emit(sd)
    return sd, emit, _gpu_tokenizer(sd, emit), c_oracle_tokenizer(sd, emit)


@pytest.mark.parametrize("kind,hmm", [("oov", True), ("long", True), ("freq", False)])
def test_benchmark_dictionary_corpora(bench_pair, kind, hmm):
    """BASELINE configs 3 / 4 / 2 with the dictionary bench.py uses, 64 MB each (several 16 MiB+ device sub-batches)."""
    from oracle import c_oracle as co
    sd, emit, tk, ora = bench_pair
    text, doc_off = synth.make_corpus(sd, kind, 64_000_000, synth.SEED_BASE + {"freq": 2, "oov": 3, "long": 4}[kind])
    t = text.numpy()
    off = doc_off.numpy().astype(np.uint64)
    _assert_same(tk.cut_batch(t, off, hmm), ora.cut_batch(t, off, hmm, co.num_procs()), t, off)


# ---- AddWord with freq < 1: suggestFreq (T:372-379, 589-614) ---------------------------------------------------
def test_add_word_suggested_frequency(small_synth):
    from oracle import py_oracle as po
    sd, emit = small_synth
    tk = _gpu_tokenizer(sd, emit)
    ora = po.Tokenizer(po.PrefixDictionary.from_lines_prefix_mode(sd.lines()), po.HiddenMarkovModel(emit))
    rng = np.random.default_rng(41)
    words = [w.decode() for w in sd.words if all(0x4E00 <= ord(c) <= 0x9FA5 for c in w.decode())]
    for trial in range(12):
        new = "".join(words[int(i)] for i in rng.integers(0, len(words), int(rng.integers(2, 4))))
        if trial % 4 == 3:
            new = words[int(rng.integers(0, len(words)))]   # an existing word: its count may only grow
        nb = new.encode()
        assert tk.suggest_freq(new) == ora.pd.suggest_freq(nb, [nb[s:e] for s, e, _ in ora.cut(nb, False)])
        tk.add_word(new, 0)
        ora.add_word(new, 0)
        assert tk.lookup(new) == ora.pd.term_freq[new.encode()] and tk.size == ora.pd.size
        ctx = words[int(rng.integers(0, len(words)))] + new + "，" + new + words[int(rng.integers(0, len(words)))]
        for hmm in (False, True):
            assert tk.cut(ctx, hmm) == ora.cut_strings(ctx, hmm)


# ---- jb_cut_device from two streams at once: one workspace, serialised on the device ---------------------------
def test_device_api_two_streams(medium_pair):
    import torch
    sd, emit, tk, ora = medium_pair
    outs = []
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    wants = []
    for i, s in enumerate(streams * 3):       # six calls, alternating streams, different sizes (the workspace grows once)
        text, doc_off = synth.make_corpus(sd, "oov" if i % 2 else "freq", 400_000 + 300_000 * (i % 3), synth.SEED_BASE + 80 + i)
        t = text.numpy()
        off = doc_off.numpy().astype(np.uint64)
        wants.append(ora.cut_batch(t, off, True, 8))
        dt, ddo = text.cuda(), doc_off.cuda()
        cap = len(wants[-1][0]) + 10
        bufs = (torch.zeros(cap, dtype=torch.int32, device="cuda"), torch.zeros(cap, dtype=torch.int32, device="cuda"),
                torch.zeros(off.size, dtype=torch.int64, device="cuda"), torch.zeros(2, dtype=torch.int64, device="cuda"), dt, ddo)
        outs.append(bufs)
    torch.cuda.synchronize()
    for i, s in enumerate(streams * 3):       # no synchronisation between the calls
        d_start, d_end, d_dto, d_nt, dt, ddo = outs[i]
        with torch.cuda.stream(s):
            tk.cut_device(dt, ddo, True, d_start, d_end, d_dto, d_nt, stream=s)
    torch.cuda.synchronize()
    for (d_start, d_end, d_dto, d_nt, _, _), (os_, oe, _, od) in zip(outs, wants):
        assert d_nt.tolist() == [len(os_), 0]
        assert np.array_equal(d_start[:len(os_)].cpu().numpy().astype(np.uint32), os_)
        assert np.array_equal(d_end[:len(os_)].cpu().numpy().astype(np.uint32), oe)
        assert np.array_equal(d_dto.cpu().numpy().astype(np.uint64), od)


# ---- the bitmap result format, pageable input, many short strings ----------------------------------------------
def _bits_equal_arrays(tk, ora, t, off, hmm, nthreads=4):
    want = ora.cut_batch(t, off, hmm, 8)
    with tk.cut_batch_bits(t, off, hmm) as r:
        assert r.n_tokens == len(want[0])
        assert np.array_equal(r.doc_tok_off, want[3])
        # the bitmaps themselves: popcounts, and every start / end bit where the oracle says
        abs_s = np.repeat(off[:-1], np.diff(want[3]).astype(np.int64)) + want[0]
        abs_e = np.repeat(off[:-1], np.diff(want[3]).astype(np.int64)) + want[1] - 1
        sb = np.unpackbits(r.start_bits.view(np.uint8), bitorder="little")
        eb = np.unpackbits(r.end_bits.view(np.uint8), bitorder="little")
        assert sb.sum() == len(want[0]) and eb.sum() == len(want[0])
        assert np.array_equal(np.nonzero(sb)[0].astype(np.uint64), abs_s - off[0])
        assert np.array_equal(np.nonzero(eb)[0].astype(np.uint64), abs_e - off[0])
        for nt in (1, nthreads):
            st, en = r.expand(nt)
            assert np.array_equal(st, want[0]) and np.array_equal(en, want[1])


@pytest.mark.parametrize("hmm", [False, True])
def test_bitmap_result_format(synth_pair, hmm):
    sd, emit, _, ora = synth_pair
    tk = _gpu_tokenizer(sd, emit, 1, max_batch_bytes=50_000)   # many sub-batches starting at every bit offset of a word
    text, doc_off = synth.make_corpus(sd, "oov", 600_000, synth.SEED_BASE + 61)
    t = text.numpy()
    rng = np.random.default_rng(6)
    cuts = np.unique(np.concatenate([[0, t.size], rng.integers(0, t.size, 300)])).astype(np.uint64)
    off = np.concatenate([cuts[:40], cuts[39:40], cuts[40:]])   # an empty document too
    _bits_equal_arrays(tk, ora, t, off, hmm)
    _bits_equal_arrays(tk, ora, t, off[5:-7], hmm)             # doc_off[0] != 0: bit 0 = first byte of the first document
    rng2 = np.random.default_rng(7)
    docs = fuzz_docs(sd, rng2, n_docs=300, max_len=60) + [b"", b"x", b""]
    ft, foff = pack_docs(docs)
    _bits_equal_arrays(tk, ora, ft, foff, hmm)
    e = np.zeros(1, np.uint64)
    with tk.cut_batch_bits(b"", e, hmm) as r:                   # no documents at all
        assert r.n_tokens == 0 and r.doc_tok_off.tolist() == [0]


def test_bitmap_result_two_tokenizers_one_device(synth_pair):
    """jb_cut_batch_multi with two tokenizers (here both on device 0; one per GPU on a multi-GPU box): two host threads,
    two pipelines, one shared bitmap result."""
    sd, emit, tk, ora = synth_pair
    tk2 = _gpu_tokenizer(sd, emit)
    tk3 = _gpu_tokenizer(sd, emit)
    text, doc_off = synth.make_corpus(sd, "freq", 1_500_000, synth.SEED_BASE + 62)
    t, off = text.numpy(), doc_off.numpy().astype(np.uint64)
    want = ora.cut_batch(t, off, True, 8)
    with tk.cut_batch_bits(t, off, True, others=[tk2, tk3]) as r:
        st, en = r.expand(3)
        assert np.array_equal(r.doc_tok_off, want[3]) and np.array_equal(st, want[0]) and np.array_equal(en, want[1])


def test_device_bits_api(medium_pair):
    import torch
    sd, emit, tk, ora = medium_pair
    text, doc_off = synth.make_corpus(sd, "oov", 1_000_000, synth.SEED_BASE + 71)
    t, off = text.numpy(), doc_off.numpy().astype(np.uint64)
    os_, oe, _, od = ora.cut_batch(t, off, True, 8)
    dt, ddo = text.cuda(), doc_off.cuda()
    nw = t.size // 32 + 8
    sb = torch.full((nw,), -1, dtype=torch.int32, device="cuda")   # (garbage in: the call clears them)
    eb = torch.full((nw,), -1, dtype=torch.int32, device="cuda")
    d_dto = torch.zeros(off.size, dtype=torch.int64, device="cuda")
    d_nt = torch.zeros(2, dtype=torch.int64, device="cuda")
    tk.cut_device_bits(dt, ddo, True, sb, eb, d_dto, d_nt)
    torch.cuda.synchronize()
    assert d_nt.tolist() == [len(os_), 0]
    assert np.array_equal(d_dto.cpu().numpy().astype(np.uint64), od)
    s_pos = np.nonzero(np.unpackbits(sb.cpu().numpy().view(np.uint8), bitorder="little")[: t.size])[0]
    e_pos = np.nonzero(np.unpackbits(eb.cpu().numpy().view(np.uint8), bitorder="little")[: t.size])[0]
    base = np.repeat(off[:-1], np.diff(od).astype(np.int64))
    assert np.array_equal(s_pos.astype(np.uint64), base + os_) and np.array_equal(e_pos.astype(np.uint64), base + oe - 1)


def test_pageable_and_pinned_input_agree(medium_pair):
    import torch
    sd, emit, tk, ora = medium_pair
    text, doc_off = synth.make_corpus(sd, "oov", 40_000_000, synth.SEED_BASE + 72)   # two 16 MiB+ sub-batches
    off = doc_off.numpy().astype(np.uint64)
    pinned = text.pin_memory()
    pageable = np.array(text.numpy(), copy=True)
    a = tk.cut_batch(pinned.numpy(), off, True)
    b = tk.cut_batch(pageable, off, True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    _assert_same(b, ora.cut_batch(pageable, off, True, 16), pageable, off)


def test_ten_thousand_short_strings_in_one_batch(synth_pair):
    """The batch call a Go caller with many short strings uses (CutBatch in go/tokenizer.go): 10k strings, one device batch."""
    sd, emit, tk, ora = synth_pair
    rng = np.random.default_rng(8)
    words = [w.decode() for w in sd.words]
    texts = []
    for i in range(10_000):
        k = int(rng.integers(0, 12))
        s = "".join(words[int(j)] if rng.random() < 0.8 else ["，", " a1 ", "。", "龥"][int(rng.integers(0, 4))] for j in rng.integers(0, len(words), k))
        texts.append(s)
    got = tk.cut_many(texts, True)
    for i in rng.integers(0, len(texts), 400).tolist() + [0, len(texts) - 1]:
        assert got[i] == ora.cut_strings(texts[i], True), texts[i]
    assert sum(len(g) for g in got) == sum(len(ora.cut_strings(s, True)) for s in texts[:2000]) + sum(len(g) for g in got[2000:])


# ---- cached table image --------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", [0, 1])
def test_cached_table_image_constructor(tmp_path, small_synth, kind):
    """jb_tokenizer_create_cached: built + written on the first call, read back on the second; both cut like the
    tokenizer built the ordinary way (and like the oracle in the same loader mode)."""
    from jieba_go_b200.tokenizer import Tokenizer
    sd, emit = small_synth
    dp, ep, ip = tmp_path / "dict.txt", tmp_path / "prob_emit.json", tmp_path / "tables.img"
    dp.write_bytes(sd.dict_txt())
    ep.write_bytes(synth.emit_json(emit))
    a = Tokenizer.from_files_cached(dp, kind, ep, ip)
    b = Tokenizer.from_files_cached(dp, kind, ep, ip)
    assert not a.from_cache and b.from_cache
    ora = c_oracle_tokenizer(sd, emit, kind)
    text, doc_off = synth.make_corpus(sd, "oov", 400_000, synth.SEED_BASE + 90)
    t, off = text.numpy(), doc_off.numpy().astype(np.uint64)
    for hmm in (False, True):
        want = ora.cut_batch(t, off, hmm, 4)
        _assert_same(a.cut_batch(t, off, hmm), want, t, off)
        _assert_same(b.cut_batch(t, off, hmm), want, t, off)
    with pytest.raises(Exception):
        b.add_word("甲乙", 5)   # read-only: no host dictionary


def test_host_pool_limit_and_numa_helper(synth_pair):
    """jb_host_pool_limit releases the pinned result pool; jb_bind_thread_to_device is best effort (single-NUMA VMs: -1)."""
    import os
    from jieba_go_b200 import _capi
    sd, emit, tk, ora = synth_pair
    L = _capi.lib()
    text, doc_off = synth.make_corpus(sd, "freq", 1_000_000, synth.SEED_BASE + 95)
    t, off = text.numpy(), doc_off.numpy().astype(np.uint64)
    tk.cut_batch(t, off, False)                 # the freed result goes to the pool ...
    assert L.jb_host_pool_limit(1 << 40) > 0    # ... and is still held
    assert L.jb_host_pool_limit(0) == 0         # released
    assert L.jb_host_pool_limit(4 << 30) == 0
    _assert_same(tk.cut_batch(t, off, False), ora.cut_batch(t, off, False, 4), t, off)   # and the pool fills again
    before = os.sched_getaffinity(0)
    node = L.jb_bind_thread_to_device(0)
    assert node >= -1 and len(os.sched_getaffinity(0)) >= 1
    os.sched_setaffinity(0, before)
