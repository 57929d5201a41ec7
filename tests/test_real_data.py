"""The reference's real-data golden vectors (tokenizer_test.go:36-47 TestCut = BASELINE config 1,
:88-126 TestBuildDAG, :275-285 TestCutDag, :291-300 TestLoadHMM, :308-317 TestViterbi, :467-473
TestBuildPrefixDictFromScratch) against both oracles and -- marked gpu -- the CUDA path through the C ABI.

They need dict.txt / prefix_dictionary.gob / prob_emit.json with the upstream sha256 (tests/realdata.py);
without them every test here SKIPS and says why.  test_locator_* always run."""
import json
import os

import numpy as np
import pytest

import kat_vectors as kv
import realdata

JIEBA_DICT_SIZE = 60_101_967  # T:454


def _need():
    paths, why = realdata.locate()
    if paths is None:
        pytest.skip("real jieba data not available: " + why)
    return paths


# ---- the locator itself ---------------------------------------------------------------------------
def test_locator_refuses_lfs_stubs(tmp_path, monkeypatch):
    for name in kv.REAL_SHA256:
        (tmp_path / name).write_bytes(b"version https://git-lfs.github.com/spec/v1\noid sha256:0\nsize 1\n")
    monkeypatch.setenv("JIEBA_DATA_DIR", str(tmp_path))
    monkeypatch.setattr(realdata, "_cache", None)
    paths, why = realdata.locate()
    assert paths is None and "Git-LFS pointer stub" in why
    monkeypatch.setattr(realdata, "_cache", None)


def test_locator_reports_missing_files(tmp_path, monkeypatch):
    monkeypatch.setenv("JIEBA_DATA_DIR", str(tmp_path))
    monkeypatch.setattr(realdata, "_cache", None)
    paths, why = realdata.locate()
    assert paths is None and "missing" in why
    monkeypatch.setattr(realdata, "_cache", None)


def test_reference_checkout_holds_stubs_only():
    """Documents F1 where the reference checkout is present (this container; not the GPU box)."""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("no reference checkout here")
    for name in kv.REAL_SHA256:
        assert os.path.getsize(os.path.join(ref, name)) < 1024


# ---- oracles ----------------------------------------------------------------------------------------
def build_oracles(paths):
    from oracle import c_oracle as co
    from oracle import py_oracle as po
    emit = json.load(open(paths["prob_emit.json"], encoding="utf-8"))
    gob = po.read_gob_map_string_int(open(paths["prefix_dictionary.gob"], "rb").read())
    # NewJiebaTokenizer (T:69-75): the gob's map as it is, size = the literal (T:454)
    ppd = po.PrefixDictionary()
    ppd.term_freq = dict(gob)
    ppd.size = JIEBA_DICT_SIZE
    ptk = po.Tokenizer(ppd, po.HiddenMarkovModel(emit))
    cpd = co.Dict()
    for k, v in gob.items():
        cpd.set_raw(k, v)
    cpd.size = JIEBA_DICT_SIZE
    ctk = co.Tokenizer(cpd, co.Hmm(emit))
    return paths, emit, gob, ptk, ctk


@pytest.fixture(scope="module")
def real_oracles():
    return build_oracles(_need())


@pytest.mark.parametrize("text,want,hmm", kv.TEST_CUT)
def test_cut_oracles(real_oracles, text, want, hmm):  # TestCut, tokenizer_test.go:36-47
    _, _, _, ptk, ctk = real_oracles
    assert ptk.cut_strings(text, hmm) == want
    assert ctk.cut_strings(text, hmm) == want


@pytest.mark.parametrize("text,want", kv.BUILD_DAG)
def test_build_dag_oracles(real_oracles, text, want):  # TestBuildDAG, tokenizer_test.go:88-126
    _, _, _, ptk, ctk = real_oracles
    runes = [ord(c) for c in text]
    assert ptk.pd.build_dag(runes) == want
    assert ctk.build_dag(text) == want


def test_cut_dag_oracles(real_oracles):  # TestCutDag, tokenizer_test.go:275-285 (= the HMM-off rows of TestCut)
    _, _, _, ptk, ctk = real_oracles
    for text, want, hmm in kv.TEST_CUT[:3:2]:
        assert not hmm
        assert ptk.cut_strings(text, False) == want and ctk.cut_strings(text, False) == want


def test_load_hmm(real_oracles):  # TestLoadHMM, tokenizer_test.go:291-300
    _, emit, _, ptk, _ = real_oracles
    for s, v in kv.LOAD_HMM.items():
        assert emit[s]["一"] == v
        assert ptk.hmm.emit_p[s][ord("一")] == v


@pytest.mark.parametrize("text,want", kv.VITERBI)
def test_viterbi_oracles(real_oracles, text, want):  # TestViterbi, tokenizer_test.go:308-317
    _, _, _, ptk, ctk = real_oracles
    assert ptk.hmm.viterbi([ord(c) for c in text]) == want
    assert ctk.hmm.viterbi(text) == want


def test_build_prefix_dict_from_scratch(real_oracles):  # tokenizer_test.go:467-473 (one-directional, :634-641)
    paths, _, gob, _, _ = real_oracles
    from oracle import py_oracle as po
    fd = po.PrefixDictionary.from_lines_prefix_mode(po.split_dict_lines(open(paths["dict.txt"], "rb").read()))
    for k, v in gob.items():
        assert fd.term_freq.get(k, 0) == v, k


def test_product_gob_reader_on_the_real_file(real_oracles):
    """jb_dict_load_gob_file against the oracle's independent gob reader (CPU: no compute entry point)."""
    import ctypes as C
    from jieba_go_b200 import _capi
    paths, _, gob, _, _ = real_oracles
    L = _capi.lib()
    db = C.c_void_p()
    _capi.check(L.jb_dict_load_gob_file(paths["prefix_dictionary.gob"].encode(), C.byref(db)), "jb_dict_load_gob_file")
    try:
        dd = _capi.DictDesc()
        L.jb_dict_buf_desc(db, C.byref(dd))
        assert dd.n == len(gob)
        rng = np.random.default_rng(1)
        keys = list(gob)
        for i in rng.integers(0, len(keys), 5000).tolist():
            v = C.c_int64()
            assert L.jb_dict_buf_lookup(db, keys[i], len(keys[i]), C.byref(v)) == 1 and v.value == gob[keys[i]]
    finally:
        L.jb_dict_buf_free(db)


# ---- the CUDA path ----------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def real_gpu(real_oracles):
    from jieba_go_b200.tokenizer import Tokenizer
    paths = real_oracles[0]
    return Tokenizer.new_jieba_tokenizer(paths["prefix_dictionary.gob"], paths["prob_emit.json"])


@pytest.mark.gpu
@pytest.mark.parametrize("text,want,hmm", kv.TEST_CUT)
def test_cut_gpu(real_gpu, text, want, hmm):  # TestCut = BASELINE.json configs[0]
    assert real_gpu.cut(text, hmm) == want


@pytest.mark.gpu
def test_new_tokenizer_file_mode_gpu(real_oracles):
    """NewTokenizer("dict.txt") (T:61-67, file mode: no prefix keys) against the oracle in the same mode."""
    from jieba_go_b200.tokenizer import Tokenizer
    from oracle import py_oracle as po
    paths, emit = real_oracles[0], real_oracles[1]
    tk = Tokenizer.new_tokenizer(paths["dict.txt"], paths["prob_emit.json"])
    fd = po.PrefixDictionary.from_lines_file_mode(po.split_dict_lines(open(paths["dict.txt"], "rb").read()))
    ora = po.Tokenizer(fd, po.HiddenMarkovModel(emit))
    for text, _, hmm in kv.TEST_CUT:
        assert tk.cut(text, hmm) == ora.cut_strings(text, hmm)


@pytest.mark.gpu
def test_route_values_real_dictionary_gpu(real_oracles, real_gpu):
    _, _, _, _, ctk = real_oracles
    han = "我昨天去上海交通大學與老師討論量子力學今天天氣很好这一刹那的撙近".encode()
    ge, gp = real_gpu.debug_route(han)
    oe, op = ctk.route(han)
    assert np.array_equal(ge, oe) and np.array_equal(gp.view(np.uint64), op.view(np.uint64))
