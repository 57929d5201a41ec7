"""tests/golden/cut_golden.json (frozen outputs of the literal Python restatement, see tests/golden/make_golden.py)
against both oracles as they are now and -- marked gpu -- the CUDA path through the C ABI."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    g = json.load(open(os.path.join(HERE, "golden", "cut_golden.json"), encoding="utf-8"))
    g["documents"] = [bytes.fromhex(h) for h in g["documents_hex"]]
    return g


def _emit():
    from golden.make_golden import kat_emit
    return kat_emit()


@pytest.mark.parametrize("case", range(4))
def test_oracles_match_the_frozen_outputs(golden, case):
    from oracle import c_oracle as co
    from oracle import py_oracle as po
    c = golden["cases"][case]
    lines = golden["dictionary_lines"]
    pd = po.PrefixDictionary.from_lines_prefix_mode(lines) if c["mode"] == 1 else po.PrefixDictionary.from_lines_file_mode(lines)
    ptk = po.Tokenizer(pd, po.HiddenMarkovModel(_emit()))
    ctk = co.Tokenizer(co.Dict.from_lines(lines, c["mode"]), co.Hmm(_emit()))
    for d, want in zip(golden["documents"], c["tokens"]):
        want = [(s, e, bool(f)) for s, e, f in want]
        assert ptk.cut(d, c["hmm"]) == want
        assert ctk.cut(d, c["hmm"]) == want


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(4))
def test_gpu_matches_the_frozen_outputs(golden, case):
    from jieba_go_b200.tokenizer import Tokenizer
    c = golden["cases"][case]
    data = "\n".join(golden["dictionary_lines"]).encode() + b"\n"
    for general in (False, True):
        tk = Tokenizer.from_dict_text(data, c["mode"], _emit())
        tk.set_general_only(general)
        docs = golden["documents"]
        off = np.zeros(len(docs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(d) for d in docs])
        st, en, dto = tk.cut_batch(b"".join(docs), off, c["hmm"])
        for i, (d, want) in enumerate(zip(docs, c["tokens"])):
            lo, hi = int(dto[i]), int(dto[i + 1])
            got = [[int(s), int(e), int(e - s == 1 and d[s] >= 0x80)] for s, e in zip(st[lo:hi], en[lo:hi])]
            assert got == want, (i, d)


# ---- long single-block documents (500..9,000 runes), pinned by token count + sha256 -------------------------------
def _long_docs(golden):
    import hashlib
    from golden.make_golden import long_documents
    docs = long_documents()
    assert hashlib.sha256(b"".join(docs)).hexdigest() == golden["long_documents_sha256"], "the seeded long documents changed"
    return docs


@pytest.mark.parametrize("case", range(4))
def test_c_oracle_matches_the_frozen_long_blocks(golden, case):
    from golden.make_golden import token_digest
    from oracle import c_oracle as co
    c = golden["long_cases"][case]
    ctk = co.Tokenizer(co.Dict.from_lines(golden["dictionary_lines"], c["mode"]), co.Hmm(_emit()))
    for d, n, h in zip(_long_docs(golden), c["n_tokens"], c["sha256"]):
        t = ctk.cut(d, c["hmm"])
        assert len(t) == n and token_digest(t) == h


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(4))
def test_gpu_matches_the_frozen_long_blocks(golden, case):
    """Blocks of 512 runes or more take the segmented walk (k_land / k_chain / k_emit over segments / k_runs)."""
    import hashlib
    from jieba_go_b200.tokenizer import Tokenizer
    c = golden["long_cases"][case]
    data = "\n".join(golden["dictionary_lines"]).encode() + b"\n"
    docs = _long_docs(golden)
    off = np.zeros(len(docs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(d) for d in docs])
    for general in (False, True):
        tk = Tokenizer.from_dict_text(data, c["mode"], _emit())
        tk.set_general_only(general)
        st, en, dto = tk.cut_batch(b"".join(docs), off, c["hmm"])
        for i, (n, h) in enumerate(zip(c["n_tokens"], c["sha256"])):
            lo, hi = int(dto[i]), int(dto[i + 1])
            pairs = np.stack([st[lo:hi], en[lo:hi]], axis=1).astype("<u4")
            assert hi - lo == n and hashlib.sha256(pairs.tobytes()).hexdigest() == h, (i, len(docs[i]) // 3, general)
