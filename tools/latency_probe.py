"""Per-call latency of the drop-in call on the reference's benchmark sentence (tokenizer_test.go:526-533:
Cut("我昨天去上海交通大學與老師討論量子力學", true), 57 bytes, published 30,726 ns/op on an i5-9400), and the
throughput of the batch call on many such strings.  Prints one JSON line.

    python tools/latency_probe.py            # small path (captured CUDA graph) on
    JB_NO_SMALL=1 python tools/latency_probe.py   # every call through the ordinary pipeline"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from jieba_go_b200 import synth
    from jieba_go_b200.tokenizer import Tokenizer
    sd = synth.make_dictionary(n_words=349_000, seed=synth.SEED_BASE)
    emit = synth.make_emit(sd)
    tk = Tokenizer.from_dict_text(sd.dict_txt(), 1, emit, device=0)
    sent = "我昨天去上海交通大學與老師討論量子力學".encode()
    off = np.array([0, len(sent)], dtype=np.uint64)
    arr = np.frombuffer(sent, dtype=np.uint8)
    for _ in range(50):
        tk.cut_batch(arr, off, True)
    ts = []
    for _ in range(2000):
        t0 = time.perf_counter()
        tk.cut_batch(arr, off, True)
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    # the batch call: 10,000 such strings in one device batch (median of 7 calls after two warm-up calls: the first
    # call of a given size allocates its pinned result buffers)
    many = [sent] * 10_000
    blob = b"".join(many)
    moff = np.arange(len(many) + 1, dtype=np.uint64) * len(sent)
    bt = []
    for it in range(9):
        t0 = time.perf_counter()
        with tk.cut_batch_bits(blob, moff, True) as r:
            nt = r.n_tokens
        bt.append(time.perf_counter() - t0)
    dt = float(np.median(bt[2:]))
    print(json.dumps({"probe": "latency", "small_path": os.environ.get("JB_NO_SMALL") is None, "bytes": len(sent),
                      "cut_us_p50": float(np.median(ts)), "cut_us_p10": float(np.percentile(ts, 10)), "cut_us_p99": float(np.percentile(ts, 99)),
                      "reference_published_us": 30.726, "batch_10k_strings_ms": dt * 1e3, "batch_us_per_string": dt * 1e6 / len(many),
                      "batch_tokens": int(nt)}))


if __name__ == "__main__":
    main()
