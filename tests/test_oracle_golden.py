"""Pins BOTH oracle restatements against the reference's data-free golden vectors
(tokenizer_test.go) and the hand-checkable micro-KATs of SURVEY.md App. D."""
import math

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import py_oracle as po
from oracle.unicode_tables import HAN_RANGES, is_han, is_space

import kat_vectors as kv


@pytest.fixture(scope="module")
def py_tk(kat_lines, kat_emit):
    return po.Tokenizer(po.PrefixDictionary.from_lines_prefix_mode(kat_lines), po.HiddenMarkovModel(kat_emit))


@pytest.fixture(scope="module")
def c_tk(kat_lines, kat_emit):
    return co.Tokenizer(co.Dict.from_lines(kat_lines, 1), co.Hmm(kat_emit))


# ---- TestSplitText (tokenizer_test.go:61-80) --------------------------------
@pytest.mark.parametrize("text,want", kv.SPLIT_TEXT)
def test_split_text_py(text, want):
    b = text.encode()
    got = [(i, b[s:e].decode(), p) for i, s, e, p in po.split_text(0, len(b), po.find_han_runs(b))]
    assert got == want


@pytest.mark.parametrize("text,want", kv.SPLIT_TEXT)
def test_split_text_c(c_tk, text, want):
    assert c_tk.split_text(text) == want


def test_split_text_empty(c_tk):
    # T:166-168: no marks => one (empty) non-Han block
    assert po.split_text(0, 0, []) == [(0, 0, 0, False)]
    assert c_tk.split_text("") == [(0, "", False)]


# ---- TestMaxIndexProba (tokenizer_test.go:136-176) --------------------------
@pytest.mark.parametrize("cands,want_idx,want_p", kv.MAX_INDEX_PROBA)
def test_max_index_proba(cands, want_idx, want_p):
    assert po.max_index_proba(cands) == (want_idx, want_p)
    assert co.max_index_proba(cands) == (want_idx, want_p)


def test_max_index_proba_is_not_argmax():
    c = [(1, -3.47), (2, -8.99), (3, -6.91)]
    assert po.max_index_proba(c)[0] == 3
    assert co.max_index_proba(c)[0] == 3
    minf = [(1, -math.inf), (2, -math.inf)]
    assert po.max_index_proba(minf)[0] == 2 and co.max_index_proba(minf)[0] == 2
    assert co.max_index_proba([(7, -math.inf)]) == (7, -math.inf)


# ---- TestFindDagPath (tokenizer_test.go:178-270) ----------------------------
@pytest.mark.parametrize("n,dag_proba,want", kv.FIND_DAG_PATH)
def test_find_dag_path(n, dag_proba, want):
    assert po.find_dag_path(n, dag_proba) == want


# ---- TestStateTransitionRoute (tokenizer_test.go:322-345) -------------------
@pytest.mark.parametrize("now,want_from", kv.STATE_ROUTE)
def test_state_transition_route(now, want_from):
    hs = {0: dict(B=1.1, M=1.1, E=1.1, S=1.1), 1: dict(B=1.1, M=1.1, E=1.1, S=1.1)}
    assert po.HiddenMarkovModel({}).state_transition_route(2, now, hs)[0] == want_from
    assert co.Hmm().state_transition_route(hs[1], now)[0] == want_from


def test_state_transition_route_none():
    m = po.MIN_FLOAT
    hs = {0: dict(B=m, M=m, E=m, S=m)}
    assert po.HiddenMarkovModel({}).state_transition_route(1, "B", hs) == ("", m)
    assert co.Hmm().state_transition_route(hs[0], "B") == ("", m)


# ---- TestCutHMM (tokenizer_test.go:347-365) ---------------------------------
@pytest.mark.parametrize("text,path,want", kv.CUT_HMM)
def test_cut_hmm(text, path, want):
    assert [text[a:b] for a, b in po.cut_hmm(len(text), path)] == want


def test_cut_hmm_short_path_drops_tail():
    assert po.cut_hmm(5, ["S"]) == [(0, 1)]


# ---- TestCutNonZh (tokenizer_test.go:367-384) -------------------------------
@pytest.mark.parametrize("text,want", kv.CUT_NON_ZH)
def test_cut_non_zh(py_tk, c_tk, text, want):
    b = text.encode()
    assert po.materialise(b, py_tk.cut_non_zh(b, 0, len(b))) == want
    assert c_tk.cut_strings(text, False) == want


# ---- TestBuildPrefixDict (tokenizer_test.go:431-465) ------------------------
def test_build_prefix_dict():
    pd = po.PrefixDictionary.from_lines_prefix_mode(kv.BUILD_PREFIX_DICT_INPUT)
    assert {k.decode(): v for k, v in pd.term_freq.items()} == kv.BUILD_PREFIX_DICT_WANT
    cd = co.Dict.from_lines(kv.BUILD_PREFIX_DICT_INPUT, 1)
    assert len(cd) == len(kv.BUILD_PREFIX_DICT_WANT)
    for k, v in kv.BUILD_PREFIX_DICT_WANT.items():
        assert cd.lookup(k) == v
    assert cd.size == pd.size == 3 + 3 + 3 + 3 + 3 + 4986


def test_file_mode_first_duplicate_wins():
    lines = ["今天 10 x", "天氣 3", "今天 99 y"]
    pd = po.PrefixDictionary.from_lines_file_mode(lines)
    assert pd.term_freq == {"今天".encode(): 10, "天氣".encode(): 3} and pd.size == 13  # T:419-423
    cd = co.Dict.from_lines(lines, 0)
    assert cd.lookup("今天") == 10 and cd.lookup("今") is None and cd.size == 13
    pd1 = po.PrefixDictionary.from_lines_prefix_mode(lines)
    assert pd1.term_freq["今天".encode()] == 99 and pd1.size == 112  # T:350-351: last wins, all counted
    cd1 = co.Dict.from_lines(lines, 1)
    assert cd1.lookup("今天") == 99 and cd1.lookup("今") == 0 and cd1.size == 112


# ---- TestAddWord (tokenizer_test.go:475-497): addTerm only ------------------
def test_add_term():
    pd = po.PrefixDictionary()
    cd = co.Dict()
    for term, freq in {"左和右": 20, "上和下": 80}.items():
        pd.add_term(term, freq)
        cd.add_term(term, freq)
    assert pd.term_freq["左和右".encode()] == 20 and pd.size == 100
    assert cd.lookup("上和下") == 80 and cd.size == 100


# ---- micro-KATs ---------------------------------------------
@pytest.mark.parametrize("text,off,on", kv.KATS)
def test_kats(py_tk, c_tk, text, off, on):
    assert py_tk.cut_strings(text, False) == off
    assert py_tk.cut_strings(text, True) == on
    assert c_tk.cut_strings(text, False) == off
    assert c_tk.cut_strings(text, True) == on


def test_kat5_selector(kat_emit):
    for tk in (po.Tokenizer(po.PrefixDictionary.from_lines_prefix_mode(kv.KAT5_LINES), po.HiddenMarkovModel(kat_emit)),
               co.Tokenizer(co.Dict.from_lines(kv.KAT5_LINES, 1), co.Hmm(kat_emit))):
        assert tk.cut_strings(kv.KAT5[0], False) == kv.KAT5[1]


def test_invalid_utf8_tokens(py_tk, c_tk):
    # Go `range` semantics (T:301-305): each ill-formed byte is one U+FFFD token of width 1
    b = b"a\xff\xe4\xb8 \xe4\xb9\x99\x80z"
    want = [(0, 1, False), (1, 2, True), (2, 3, True), (3, 4, True), (5, 8, False), (8, 9, True), (9, 10, False)]
    assert py_tk.cut(b, False) == want
    assert c_tk.cut(b, False) == want
    assert c_tk.cut_strings(b, False) == ["a", "�", "�", "�", "乙", "�", "z"]


# ---- Go stdlib restatements -------------------------------------------------
def test_go_log_restatements_agree():
    assert po.go_log(60101967.0).hex() == "0x1.1e95b8bb84672p+4"  # SURVEY App. E check value
    rng = np.random.default_rng(1)
    xs = np.concatenate([np.arange(1, 20000), rng.integers(1, 2 ** 31, 20000)]).astype(np.float64)
    for x in xs.tolist():
        assert po.go_log(x) == co.go_log(x)
    assert co.go_log(0.0) == -math.inf and co.go_log(1.0) == 0.0
    # it is NOT glibc's log: differs by 1 ulp on ~1 % of integers (e.g. 3)
    assert co.go_log(3.0) != math.log(3.0)


def test_unicode_tables():
    import regex
    han = regex.compile(r"\p{Script=Han}")
    # every range of both committed tables is Han in the (newer) regex-module table
    for ver in (13, 15):
        for lo, hi in HAN_RANGES[ver]:
            for cp in {lo, hi, (lo + hi) // 2}:
                assert han.fullmatch(chr(cp)), hex(cp)
    for ch in "中文々〇㐀":
        assert is_han(ord(ch))
    for ch in "，。、『』ステ번a1 ":
        assert not is_han(ord(ch))
    sp = regex.compile(r"\p{White_Space}")
    for cp in list(range(0, 0x3100)) + [0xFEFF, 0x1680]:
        assert is_space(cp) == bool(sp.fullmatch(chr(cp))), hex(cp)
