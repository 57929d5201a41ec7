"""Seeded synthetic workloads for the Cut hot path.

The reference's dict.txt / prefix_dictionary.gob / prob_emit.json are Git-LFS
stubs in /root/reference (132 bytes each), so tests and benchmarks run on a
synthetic jieba-format dictionary, emission table and corpora generated here.
Nothing in this file is on the product path; it only produces inputs.

  make_dictionary()  -> SynthDict   (dict.txt lines: "word freq [pos]")
  make_emit()        -> {"B": {rune: logp}, "M": ..., "E": ..., "S": ...}
  make_corpus()      -> (text uint8 tensor, doc_off int64 tensor) for the
                        BASELINE.json configs:
      "freq"  config 2: words sampled i.i.d. by frequency, sentences of
              U{4..32} words joined by one of ，。！？； ; every 8th separator
              is an ASCII group like " a1b2 "; docs = 64 sentences + "\\n".
      "oov"   config 3/5: as "freq" with 30 % of Han runes replaced by a
              uniform rune from U+4E00..U+9FA5.
      "long"  config 4: docs = one 10,000-rune Han-only block (10 % random
              runes) + "\\n".
Seed convention: seed = 0x6A696562 + config_id.
"""
import json
from dataclasses import dataclass

import numpy as np
import torch

SEED_BASE = 0x6A696562
HAN_LO, HAN_HI = 0x4E00, 0x9FA5  # synthetic text stays inside the Han range common to all Unicode versions
N_HAN = HAN_HI - HAN_LO + 1      # 20,902

_POS = [b"n", b"v", b"a", b"ns", b"nz", b"nr", b"d", b"m", b"i", b"l"]


@dataclass
class SynthDict:
    words: list            # list[bytes], dict.txt order (may contain duplicates)
    freqs: np.ndarray      # int64, same order
    pos: list              # list[bytes | None]
    rune_rank: np.ndarray  # code points, most popular first

    def dict_txt(self) -> bytes:
        out = []
        for w, f, p in zip(self.words, self.freqs.tolist(), self.pos):
            out.append(w + b" " + str(f).encode() + (b" " + p if p is not None else b""))
        return b"\n".join(out) + b"\n"

    def lines(self):
        return self.dict_txt().split(b"\n")[:-1]


def _rune_popularity(n):
    r = np.arange(n, dtype=np.float64)
    p = 1.0 / np.power(r + 5.0, 1.3)
    return p / p.sum()


def make_dictionary(n_words: int = 349_000, seed: int = SEED_BASE, total_freq: float = 6.0e7,
                    max_len: int = 16) -> SynthDict:
    """Jieba-format dictionary: ~5 % single runes (the most popular ones), 40 % / 30 % / 20 %
    of length 2 / 3 / 4, the rest 5..max_len; Zipf(1) integer frequencies scaled to total_freq;
    a few mixed-script entries and duplicate lines (first-wins vs last-wins loaders differ on them)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rank = rng.permutation(N_HAN).astype(np.int64) + HAN_LO
    pop = _rune_popularity(N_HAN)
    n_single = min(max(1, int(n_words * 0.05)), N_HAN - 64)
    # leave a few popular runes without a single-rune entry so that "missing" and
    # "prefix-only (freq 0)" first runes (T:468-472) occur in frequency-sampled text too
    holes = set(rng.choice(np.arange(50, max(51, min(n_single, 4000))), size=min(12, max(1, n_single // 40)), replace=False).tolist())
    singles = [int(rank[i]) for i in range(n_single) if i not in holes]
    n_multi = n_words - len(singles)
    lens_choices = np.array([2, 3, 4] + list(range(5, max_len + 1)))
    tail = np.array([1.0 / (k - 3) ** 1.5 for k in range(5, max_len + 1)])
    tail = 0.05 * tail / tail.sum() if len(tail) else tail
    probs = np.concatenate([[0.42, 0.32, 0.21], tail])
    probs = probs / probs.sum()
    seen = set()
    multi = []
    cdf = np.cumsum(pop)
    while len(multi) < n_multi:
        need = int((n_multi - len(multi)) * 1.3) + 16
        ls = rng.choice(lens_choices, size=need, p=probs)
        tot = int(ls.sum())
        rr = np.searchsorted(cdf, rng.random(tot), side="right").clip(0, N_HAN - 1)
        cps = rank[rr]
        pos = 0
        for L in ls.tolist():
            w = tuple(cps[pos:pos + L].tolist())
            pos += L
            if w in seen:
                continue
            seen.add(w)
            multi.append(w)
            if len(multi) >= n_multi:
                break
    # frequency ranks: singles get ranks spread over the head, multi-rune words a random order
    all_words = [(c,) for c in singles] + multi
    n = len(all_words)
    order = np.empty(n, dtype=np.float64)
    order[:len(singles)] = np.arange(len(singles)) * 3.0 + rng.random(len(singles))
    order[len(singles):] = rng.random(len(multi)) * n * 1.2 + 10.0
    rk = np.argsort(np.argsort(order)).astype(np.float64)
    c = total_freq / (np.log(n) + 0.5772)
    freqs = np.maximum(1, np.floor(c / (rk + 1.0))).astype(np.int64)
    words = ["".join(map(chr, w)).encode("utf-8") for w in all_words]
    # mixed-script entries (never reachable from a Han block; they exercise the loaders)
    extra = [(b"AT&T", 3), ("B超".encode(), 3), (b"c#", 3), ("江南style".encode(), 3), ("江南".encode(), 4986)]
    for w, f in extra:
        words.append(w)
        freqs = np.append(freqs, f)
    # duplicate lines with a different count
    ndup = min(8, n // 10)
    for i in rng.choice(n, size=ndup, replace=False).tolist():
        words.append(words[i])
        freqs = np.append(freqs, int(freqs[i]) // 2 + 1)
    perm = rng.permutation(len(words))
    words = [words[i] for i in perm]
    freqs = freqs[perm]
    pos = [(_POS[int(x)] if x < len(_POS) else None) for x in rng.integers(0, len(_POS) + 2, size=len(words))]
    return SynthDict(words, freqs, pos, rank)


def make_emit(sd: SynthDict, seed: int = SEED_BASE + 100):
    """Per-state emission tables over deliberately incomplete subsets of the Han range
    (sizes modelled on jieba's prob_emit.json: B 6.9k, M 6.4k, E 7.4k, S 14.5k), values U(-12,-3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sizes = {"B": 6857, "M": 6409, "E": 7439, "S": 14519}
    emit = {}
    for s in "BMES":
        k = min(sizes[s], N_HAN)
        top = sd.rune_rank[:k]
        keep = rng.random(k) > 0.04                      # holes among popular runes
        extra = sd.rune_rank[k:][rng.random(N_HAN - k) < 0.02]
        cps = np.concatenate([top[keep], extra])
        vals = rng.uniform(-12.0, -3.0, size=len(cps))
        emit[s] = {int(c): float(v) for c, v in zip(cps.tolist(), vals.tolist())}
    return emit


def emit_json(emit) -> bytes:
    """prob_emit.json format: {"B": {"<char>": <float>, ...}, ...}."""
    obj = {s: {chr(c): v for c, v in tab.items()} for s, tab in emit.items()}
    return json.dumps(obj, ensure_ascii=False).encode("utf-8")


# ----------------------------------------------------------------------------
# corpora (torch, device-agnostic: runs on CPU in tests and on cuda in bench.py)
# ----------------------------------------------------------------------------
_PUNCT = ["，", "。", "！", "？", "；"]
_ASCII_SEPS = [" a1b2 ", " x9 ", " GPU2024 ", " b200 ", " v1 ", " Go118 ", " sm100a ", " hbm3e "]


class _PieceTable:
    """Byte strings addressable by id: dictionary words (freq>0, Han only) then separators."""

    def __init__(self, sd: SynthDict, device):
        agg = {}
        for w, f in zip(sd.words, sd.freqs.tolist()):
            if w not in agg:
                agg[w] = f
        words, freqs = [], []
        for w, f in agg.items():
            if f > 0 and all(HAN_LO <= ord(ch) <= HAN_HI for ch in w.decode("utf-8")):
                words.append(w)
                freqs.append(f)
        self.n_words = len(words)
        seps = [p.encode("utf-8") for p in _PUNCT] + [a.encode() for a in _ASCII_SEPS] + [b"\n"]
        self.punct0 = self.n_words
        self.ascii0 = self.punct0 + len(_PUNCT)
        self.newline = self.ascii0 + len(_ASCII_SEPS)
        pieces = words + seps
        lens = np.array([len(p) for p in pieces], dtype=np.int64)
        off = np.zeros(len(pieces) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        blob = np.frombuffer(b"".join(pieces), dtype=np.uint8).copy()
        self.blob = torch.from_numpy(blob).to(device)
        self.off = torch.from_numpy(off).to(device)
        self.len = torch.from_numpy(lens).to(device)
        f = np.array(freqs, dtype=np.float64)
        self.cdf = torch.from_numpy(np.cumsum(f) / f.sum()).to(device)
        self.device = device


def _assemble(pt: _PieceTable, ids: torch.Tensor) -> torch.Tensor:
    lens = pt.len[ids]
    start = torch.cumsum(lens, 0) - lens
    total = int(lens.sum().item())
    piece_of_byte = torch.repeat_interleave(torch.arange(ids.numel(), device=ids.device), lens, output_size=total)
    within = torch.arange(total, device=ids.device) - start[piece_of_byte]
    src = pt.off[ids][piece_of_byte] + within
    return pt.blob[src]


def _replace_han(text: torch.Tensor, frac: float, gen: torch.Generator) -> torch.Tensor:
    """Replace `frac` of the 3-byte Han runes (lead 0xE4..0xE9) by a uniform rune from U+4E00..U+9FA5."""
    lead = ((text >= 0xE4) & (text <= 0xE9)).nonzero(as_tuple=True)[0]
    if lead.numel() == 0 or frac <= 0:
        return text
    pick = torch.rand(lead.numel(), generator=gen, device=text.device) < frac
    pos = lead[pick]
    cp = torch.randint(HAN_LO, HAN_HI + 1, (pos.numel(),), generator=gen, device=text.device)
    text = text.clone()
    text[pos] = (0xE0 | (cp >> 12)).to(torch.uint8)
    text[pos + 1] = (0x80 | ((cp >> 6) & 0x3F)).to(torch.uint8)
    text[pos + 2] = (0x80 | (cp & 0x3F)).to(torch.uint8)
    return text


def _chunk_freq(pt: _PieceTable, target_bytes: int, gen: torch.Generator, oov: float, sent_base: int):
    dev = pt.device
    # ~ (18 words * ~2.1 runes * 3 B) + sep  ~ 118 B / sentence ; oversample then trim to whole docs
    n_sent = max(64, int(target_bytes / 105) // 64 * 64 + 64)
    k = torch.randint(4, 33, (n_sent,), generator=gen, device=dev)
    n_words = int(k.sum().item())
    wid = torch.searchsorted(pt.cdf, torch.rand(n_words, generator=gen, device=dev, dtype=torch.float64)).clamp_(max=pt.n_words - 1)
    sidx = torch.arange(n_sent, device=dev) + sent_base
    sep = pt.punct0 + torch.randint(0, len(_PUNCT), (n_sent,), generator=gen, device=dev)
    is_ascii = (sidx % 8) == 7
    sep = torch.where(is_ascii, pt.ascii0 + (sidx // 8) % len(_ASCII_SEPS), sep)
    doc_end = (sidx % 64) == 63
    # pieces per sentence: k words + sep + optional newline
    per = k + 1 + doc_end.long()
    pstart = torch.cumsum(per, 0) - per
    total = int(per.sum().item())
    ids = torch.empty(total, dtype=torch.long, device=dev)
    sent_of_word = torch.repeat_interleave(torch.arange(n_sent, device=dev), k, output_size=n_words)
    wstart = torch.cumsum(k, 0) - k
    ids[pstart[sent_of_word] + (torch.arange(n_words, device=dev) - wstart[sent_of_word])] = wid
    ids[pstart + k] = sep
    de = doc_end.nonzero(as_tuple=True)[0]
    ids[pstart[de] + k[de] + 1] = pt.newline
    text = _assemble(pt, ids)
    if oov > 0:
        text = _replace_han(text, oov, gen)
    return text, n_sent


def make_corpus(sd: SynthDict, kind: str, nbytes: int, seed: int, device="cpu", chunk_bytes: int = 32 << 20):
    """-> (text uint8[n], doc_off int64[ndocs+1]); n is the largest whole-document size <= ~nbytes
    (at least one document)."""
    device = torch.device(device)
    pt = _PieceTable(sd, device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    parts = []
    have = 0
    if kind in ("freq", "oov"):
        oov = 0.30 if kind == "oov" else 0.0
        sent_base = 0
        while have < nbytes:
            t, ns = _chunk_freq(pt, min(chunk_bytes, nbytes - have), gen, oov, sent_base)
            sent_base += ns
            parts.append(t)
            have += t.numel()
    elif kind == "long":
        run = 10_000
        n_docs = max(1, nbytes // (run * 3 + 1))
        per_chunk = max(1, chunk_bytes // (run * 3 + 1))
        done = 0
        while done < n_docs:
            nd = min(per_chunk, n_docs - done)
            need_runes = nd * run
            n_words = int(need_runes / 1.6) + 64
            chunks, got = [], 0
            while got < need_runes * 3:
                wid = torch.searchsorted(pt.cdf, torch.rand(n_words, generator=gen, device=device, dtype=torch.float64)).clamp_(max=pt.n_words - 1)
                b = _assemble(pt, wid)
                chunks.append(b)
                got += b.numel()
            han = torch.cat(chunks)[: need_runes * 3]
            han = _replace_han(han, 0.10, gen).view(nd, run * 3)
            nl = torch.full((nd, 1), 0x0A, dtype=torch.uint8, device=device)
            parts.append(torch.cat([han, nl], dim=1).reshape(-1))
            done += nd
    else:
        raise ValueError("unknown corpus kind %r" % kind)
    text = torch.cat(parts) if len(parts) > 1 else parts[0]
    nl = (text == 0x0A).nonzero(as_tuple=True)[0]
    ends = nl + 1
    # trim to whole documents not exceeding nbytes (keep at least one)
    keep = int((ends <= max(nbytes, int(ends[0].item()))).sum().item())
    ends = ends[:keep]
    text = text[: int(ends[-1].item())]
    doc_off = torch.cat([torch.zeros(1, dtype=torch.long, device=device), ends])
    return text.contiguous(), doc_off
