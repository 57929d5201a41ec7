"""Regenerates tests/golden/cut_golden.json:  python tests/golden/make_golden.py

The reference is Go and cannot run in this image, so these are NOT outputs of the reference: they are outputs of
the literal Python restatement (oracle/py_oracle.py), frozen in history so that a later change to either oracle or
to the CUDA path that alters any token shows up as a diff against a committed file.  Inputs: the SURVEY App. D
dictionary / emissions and a seeded set of documents (words, random Han, ASCII, punctuation, other scripts,
ill-formed UTF-8, 4-byte Han), HMM off and on, both dictionary modes."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import py_oracle as po  # noqa: E402

KAT_LINES = "甲 100,甲甲 50,甲甲甲甲 7,乙 40,乙丙 30,乙丙丁 5,丙 20,丁 60,丙丁 25,戊己 9,己 3,庚 8,辛 2".split(",")


def kat_emit():
    emit = {s: {c: -5.0 - i * 0.1 for i, c in enumerate("甲乙丙丁己庚辛")} for s in "BMES"}
    emit["S"]["壬"] = -6.0
    return emit


def documents(seed=20261018, n=120):
    rng = np.random.default_rng(seed)
    alphabet = list("甲乙丙丁戊己庚辛壬癸") + ["甲甲", "乙丙丁", "丙丁", "戊己", "甲甲甲甲"]
    misc = ["，", "。", " ", "\t", "\n", "　", "a", "Z9", "+", "=", "번역", "ステ", "ＡＢ", "\U00020000", "々", "é", "€"]
    bad = [b"\xff", b"\xc0\x80", b"\xe4\xb8", b"\xed\xa0\x80", b"\x80", b"\xf0\x9f"]
    docs = []
    for _ in range(n):
        parts = []
        for _ in range(int(rng.integers(0, 40))):
            r = rng.random()
            if r < 0.7:
                parts.append(alphabet[int(rng.integers(0, len(alphabet)))].encode())
            elif r < 0.93:
                parts.append(misc[int(rng.integers(0, len(misc)))].encode())
            else:
                parts.append(bad[int(rng.integers(0, len(bad)))])
        docs.append(b"".join(parts))
    return docs


def long_documents(seed=20261019):
    """Documents made of ONE Han block each, 500 to 9,000 runes (the lengths at which the CUDA path cuts a block into
    256-rune segments, and around every segment edge), with unknown runes in runs of 1 to 40 (Viterbi runs that cross
    the segment boundaries)."""
    rng = np.random.default_rng(seed)
    words = ["甲", "乙", "丙", "丁", "己", "庚", "辛", "甲甲", "乙丙", "乙丙丁", "丙丁", "戊己", "甲甲甲甲"]
    unknown = "壬癸戊子丑寅卯"
    docs = []
    for n in (500, 511, 512, 513, 767, 768, 769, 1024, 1281, 2047, 2048, 2049, 4100, 9000):
        out, have = [], 0
        while have < n:
            if rng.random() < 0.25:
                k = int(rng.integers(1, 41)) if rng.random() < 0.15 else int(rng.integers(1, 4))
                piece = "".join(unknown[int(i)] for i in rng.integers(0, len(unknown), k))
            else:
                piece = words[int(rng.integers(0, len(words)))]
            piece = piece[: n - have]
            out.append(piece)
            have += len(piece)
        docs.append("".join(out).encode())
    return docs


def token_digest(tokens):
    """sha256 over the tokens as little-endian uint32 (start, end) pairs: long documents are pinned by a hash."""
    import hashlib
    return hashlib.sha256(np.asarray([[s, e] for s, e, _ in tokens], dtype="<u4").tobytes()).hexdigest()


def main():
    emit = kat_emit()
    out = {"dictionary_lines": KAT_LINES, "cases": []}
    docs = documents()
    for mode, name in ((1, "prefix"), (0, "file")):
        pd = po.PrefixDictionary.from_lines_prefix_mode(KAT_LINES) if mode == 1 else po.PrefixDictionary.from_lines_file_mode(KAT_LINES)
        tk = po.Tokenizer(pd, po.HiddenMarkovModel(emit))
        for hmm in (False, True):
            toks = [[[s, e, int(f)] for s, e, f in tk.cut(d, hmm)] for d in docs]
            out["cases"].append({"mode": mode, "mode_name": name, "hmm": hmm, "tokens": toks})
    out["documents_hex"] = [d.hex() for d in docs]
    # long single-block documents: regenerated from the seed by the tests, pinned here by token count + digest
    ldocs = long_documents()
    out["long_cases"] = []
    for mode, name in ((1, "prefix"), (0, "file")):
        pd = po.PrefixDictionary.from_lines_prefix_mode(KAT_LINES) if mode == 1 else po.PrefixDictionary.from_lines_file_mode(KAT_LINES)
        tk = po.Tokenizer(pd, po.HiddenMarkovModel(emit))
        for hmm in (False, True):
            toks = [tk.cut(d, hmm) for d in ldocs]
            out["long_cases"].append({"mode": mode, "hmm": hmm, "n_tokens": [len(t) for t in toks], "sha256": [token_digest(t) for t in toks]})
    import hashlib
    out["long_documents_sha256"] = hashlib.sha256(b"".join(ldocs)).hexdigest()
    with open(os.path.join(HERE, "cut_golden.json"), "w") as f:
        json.dump(out, f, ensure_ascii=False, separators=(",", ":"))
    print("wrote", sum(len(t) for c in out["cases"] for t in c["tokens"]), "tokens for", len(docs), "documents x 4 cases")


if __name__ == "__main__":
    main()
