"""include/jieba_b200.hpp -- the C++ mirror of the reference's interface (NewTokenizer / Cut / CutParallel / AddWord,
tokenizer.go:52-162, 372-379) -- compiled with g++ against libjieba_b200.so and driven by tests/cpp/host_mirror_test.cpp.
This is the compiled-language caller this image can build (go/tokenizer.go needs a Go toolchain): the same pointer
passing, the same U+FFFD materialisation, the same rebuild-and-swap AddWord.  Expected tokens come from the oracle."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

import kat_vectors as kv
from helpers import fuzz_docs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")


def _build(tmp_path):
    import jieba_go_b200.build as jb
    jb.build()  # the library must exist (no-op when it is up to date)
    libdir = os.path.join(ROOT, "jieba_go_b200")
    exe = str(tmp_path / "host_mirror_test")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
           "-L", libdir, "-ljieba_b200", "-Wl,-rpath," + libdir, "-pthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _hex(s):
    b = s.encode("utf-8") if isinstance(s, str) else bytes(s)
    return b.hex() if b else "-"


def _hexlist(tokens):
    return ",".join(_hex(t) for t in tokens) if tokens else "-"


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_header_compiles_and_fails_loudly_without_a_device(tmp_path):
    """-Wall -Wextra -Werror clean; without a CUDA device construction throws Error{JB_ECUDA}: no CPU fallback behind
    the mirror either."""
    import torch
    exe = _build(tmp_path)
    r = subprocess.run([exe, "--no-device"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 3, r.stdout
    else:
        assert r.returncode == 0, r.stdout
        assert "no usable CUDA device" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 0])
def test_cpp_mirror_against_the_oracle(tmp_path, kat_lines, kat_emit, mode):
    from oracle import py_oracle as po
    exe = _build(tmp_path)
    dict_path, emit_path, case_path = tmp_path / "dict.txt", tmp_path / "prob_emit.json", tmp_path / "cases.txt"
    dict_path.write_bytes(("\n".join(kat_lines) + "\n").encode())
    emit_path.write_text(json.dumps(kat_emit, ensure_ascii=False))
    pd = po.PrefixDictionary.from_lines_prefix_mode(kat_lines) if mode == 1 else po.PrefixDictionary.from_lines_file_mode(kat_lines)
    ora = po.Tokenizer(pd, po.HiddenMarkovModel(kat_emit))

    def expect(b, hmm):
        return [b"\xef\xbf\xbd" if f else b[s:e] for s, e, f in ora.cut(b, hmm)]

    rec = ["dict %s %d" % (dict_path, mode), "emit %s" % emit_path, "unicode 15"]
    texts = [t.encode() for t, _, _ in kv.KATS] + [t.encode() for t, _ in kv.CUT_NON_ZH] + [t.encode() for t, _ in kv.SPLIT_TEXT]
    # ill-formed UTF-8 next to Han and inside ASCII runs (Go `range` semantics, T:301-305)
    texts += [b"\xe7\x94\xb2\xff\xe4\xb9\x99", b"a\xffb \x80", b"\xe7\x94", b"\xf0\x9f\xe7\x94\xb2\xe7\x94\xb2", b"x" * 300 + "甲甲甲甲".encode() * 700]
    rng = np.random.default_rng(5)

    class _SD:  # fuzz_docs wants .words
        words = [ln.split(" ")[0].encode() for ln in kat_lines]

    texts += fuzz_docs(_SD, rng, n_docs=60, max_len=60)
    for b in texts:
        for hmm in (0, 1):
            rec.append("cut %d %s %s" % (hmm, _hex(b), _hexlist(expect(b, bool(hmm)))))
    if mode == 1:  # the hand-checked expectations of App. D hold for the prefix-mode dictionary
        for t, off, on in kv.KATS:
            assert expect(t.encode(), False) == [x.encode() for x in off] and expect(t.encode(), True) == [x.encode() for x in on]
    long_text = ("乙丙，a1 乙丙甲甲甲" * 50).encode()
    for workers, ordered in ((1, 1), (4, 1), (8, 0)):  # any block order conforms when ordered is false; ours is Cut's
        rec.append("par 1 %d %d %s %s" % (workers, ordered, _hex(long_text), _hexlist(expect(long_text, True))))
    rec.append("batch")
    # AddWord with a given count and with a suggested one (T:372-379, 589-614)
    rec.append("add %s 5000" % _hex("甲乙"))
    ora.add_word("甲乙", 5000)
    rec.append("freq %s %d" % (_hex("甲乙"), ora.pd.term_freq["甲乙".encode()]))
    rec.append("add %s 0" % _hex("丁庚辛"))
    ora.add_word("丁庚辛", 0)
    rec.append("freq %s %d" % (_hex("丁庚辛"), ora.pd.term_freq["丁庚辛".encode()]))
    for b in ["甲乙", "丁庚辛甲乙丙", "乙丁庚辛"]:
        for hmm in (0, 1):
            rec.append("cut %d %s %s" % (hmm, _hex(b), _hexlist(expect(b.encode(), bool(hmm)))))
    case_path.write_text("\n".join(rec) + "\n")
    r = subprocess.run([exe, str(case_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " 0 mismatches" in r.stdout
