"""Shared helpers for building oracle / product tokenizers from synthetic data."""
import numpy as np

from oracle import c_oracle as co
from oracle import py_oracle as po


def emit_arrays(emit):
    st = np.array(["BMES".index(s) for s in "BMES" for _ in emit[s]], dtype=np.uint8)
    ru = np.array([c for s in "BMES" for c in emit[s]], dtype=np.uint32)
    va = np.array([v for s in "BMES" for v in emit[s].values()], dtype=np.float64)
    return st, ru, va


def c_oracle_tokenizer(sd, emit, mode=1, unicode_version=15):
    pd = co.Dict.from_lines(sd.dict_txt(), mode)
    hm = co.Hmm()
    hm.set_emit_arrays(*emit_arrays(emit))
    return co.Tokenizer(pd, hm, unicode_version)


def py_oracle_tokenizer(sd, emit, mode=1, unicode_version=15):
    lines = sd.lines()
    pd = po.PrefixDictionary.from_lines_prefix_mode(lines) if mode == 1 else po.PrefixDictionary.from_lines_file_mode(lines)
    return po.Tokenizer(pd, po.HiddenMarkovModel(emit), unicode_version)


_MISC = ["，", "。", " ", "\t", "\n", "　", "a", "Z9", "+", "=", "번역", "ステ", "ＡＢ",
         "\U00020000", "\U00020001\U0002A700", "々", "〇", "", " ", " ", "é", "€"]
_BAD = [b"\xff", b"\xc0\x80", b"\xe4\xb8", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\x80", b"\xbf\xbf",
        b"\xf0\x9f", b"\xe0\x80\x80"]


def fuzz_docs(sd, rng, n_docs=40, max_len=120, supp_han=True):
    """Random documents mixing dictionary words, random Han (BMP + supplementary), ASCII, spaces,
    punctuation, other scripts and ill-formed UTF-8.  supp_han=False leaves out the 4-byte Han runes
    (one of them sends the whole device batch to the general kernels)."""
    words = sd.words
    misc = _MISC if supp_han else [m for m in _MISC if not any(ord(c) >= 0x20000 for c in m)]
    docs = []
    for _ in range(n_docs):
        parts = []
        for _ in range(int(rng.integers(0, max_len))):
            r = rng.random()
            if r < 0.55:
                parts.append(words[int(rng.integers(0, len(words)))])
            elif r < 0.75:
                parts.append(chr(int(rng.integers(0x4E00, 0x9FA6))).encode())
            elif r < 0.93:
                parts.append(misc[int(rng.integers(0, len(misc)))].encode())
            else:
                parts.append(_BAD[int(rng.integers(0, len(_BAD)))])
        docs.append(b"".join(parts))
    return docs


def pack_docs(docs):
    text = np.frombuffer(b"".join(docs), dtype=np.uint8)
    off = np.zeros(len(docs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(d) for d in docs])
    return text, off
