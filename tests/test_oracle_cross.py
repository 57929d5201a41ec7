"""The fast C restatement must agree with the literal Python restatement on randomised inputs
(both dictionary modes, HMM on/off, ill-formed UTF-8, supplementary-plane Han)."""
import numpy as np
import pytest

from jieba_go_b200 import synth

from helpers import c_oracle_tokenizer, fuzz_docs, py_oracle_tokenizer


@pytest.mark.parametrize("mode", [1, 0])
def test_c_vs_py_on_fuzz(small_synth, mode):
    sd, emit = small_synth
    ctk = c_oracle_tokenizer(sd, emit, mode)
    ptk = py_oracle_tokenizer(sd, emit, mode)
    assert len(ctk.pd) == len(ptk.pd.term_freq) and ctk.pd.size == ptk.pd.size
    rng = np.random.default_rng(7 + mode)
    docs = fuzz_docs(sd, rng, n_docs=60)
    for hmm in (False, True):
        for d in docs:
            assert ctk.cut(d, hmm) == ptk.cut(d, hmm), d
    assert ctk.hmm.route_ties == 0 and ptk.hmm.route_ties == 0


@pytest.mark.parametrize("kind", ["freq", "oov", "long"])
def test_c_vs_py_on_corpora(small_synth, kind):
    sd, emit = small_synth
    ctk = c_oracle_tokenizer(sd, emit)
    ptk = py_oracle_tokenizer(sd, emit)
    text, doc_off = synth.make_corpus(sd, kind, 40_000, synth.SEED_BASE + 20)
    t = text.numpy()
    off = doc_off.numpy().astype(np.uint64)
    for hmm in (False, True):
        s, e, f, dto = ctk.cut_batch(t, off, hmm, nthreads=3)
        assert dto[-1] == len(s)
        for d in range(len(off) - 1):
            doc = t[off[d]:off[d + 1]].tobytes()
            want = ptk.cut(doc, hmm)
            sl = slice(int(dto[d]), int(dto[d + 1]))
            got = list(zip(s[sl].tolist(), e[sl].tolist(), [bool(x) for x in f[sl]]))
            assert got == want


def test_batch_threads_agree(small_synth):
    sd, emit = small_synth
    ctk = c_oracle_tokenizer(sd, emit)
    text, doc_off = synth.make_corpus(sd, "oov", 300_000, synth.SEED_BASE + 21)
    t = text.numpy()
    off = doc_off.numpy().astype(np.uint64)
    a = ctk.cut_batch(t, off, True, 1)
    b = ctk.cut_batch(t, off, True, 5)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
