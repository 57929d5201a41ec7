"""Randomised parity soak (not part of pytest): many seeds x dictionaries x modes x HMM, GPU vs the C oracle.
python tools/soak.py [n_rounds]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from jieba_go_b200 import synth
from jieba_go_b200.tokenizer import Tokenizer
from helpers import c_oracle_tokenizer, fuzz_docs, pack_docs

n_rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
bad = 0
t0 = time.time()
for rnd in range(n_rounds):
    rng = np.random.default_rng(seed0 + rnd)
    n_words = int(rng.choice([300, 3000, 20000, 80000]))
    max_len = int(rng.choice([3, 6, 10, 16, 24]))
    sd = synth.make_dictionary(n_words=n_words, seed=synth.SEED_BASE + 500 + rnd, total_freq=float(rng.choice([1e5, 5e6, 6e7])), max_len=max_len)
    emit = synth.make_emit(sd, seed=synth.SEED_BASE + 600 + rnd)
    for mode in (1, 0):
        tk = Tokenizer.from_dict_text(sd.dict_txt(), mode, emit, max_batch_bytes=int(rng.choice([3_000_000, 1 << 27])))
        ora = c_oracle_tokenizer(sd, emit, mode)
        docs = fuzz_docs(sd, rng, n_docs=int(rng.integers(50, 1500)), max_len=int(rng.integers(5, 400)), supp_han=bool(rng.integers(0, 2)))
        text, off = pack_docs(docs)
        cases = [(text, off)]
        for kind in ("freq", "oov", "long"):
            t, d = synth.make_corpus(sd, kind, int(rng.integers(200_000, 2_500_000)), synth.SEED_BASE + 700 + rnd)
            t = t.numpy(); d = d.numpy().astype(np.uint64)
            cases.append((t, d))
            cuts = np.unique(np.concatenate([[0, t.size], rng.integers(0, t.size, int(rng.integers(1, 300)))])).astype(np.uint64)
            cases.append((t, cuts))
        for ci, (t, d) in enumerate(cases):
            for hmm in (False, True):
                g = tk.cut_batch(t, d, hmm)
                o = ora.cut_batch(t, d, hmm, 8)
                ok = np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]) and np.array_equal(g[2], o[3])
                with tk.cut_batch_bits(t, d, hmm) as r:   # the bitmap result format, expanded on the host
                    bs, be = r.expand(4)
                    ok = ok and np.array_equal(r.doc_tok_off, o[3]) and np.array_equal(bs, o[0]) and np.array_equal(be, o[1])
                if not ok:
                    bad += 1
                    print("MISMATCH round %d mode %d case %d hmm %s (n_words %d max_len %d)" % (rnd, mode, ci, hmm, n_words, max_len), flush=True)
        tk.close()
    print("round %d done (%d words, max_len %d) %.0fs" % (rnd, n_words, max_len, time.time() - t0), flush=True)
print("SOAK", "FAILED %d" % bad if bad else "OK", "rounds", n_rounds)
sys.exit(1 if bad else 0)
