#!/usr/bin/env python
"""Benchmark of the Cut hot path (BASELINE.json metric: UTF-8 MB/s segmented, % of HBM roofline).

    python bench.py --gpus N --steps K --warmup W [--config 2|3|4|5] [--bytes B]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port; Go is absent)

A step = one pass of the whole Cut pipeline over one synthetic batch.  At N=1 the default
workload is BASELINE.json configs[1]: 1e9 bytes sampled from the (synthetic, jieba-format)
dictionary by frequency, HMM off.  With N>1 (torchrun, one rank per GPU) every rank cuts its own
batch of the same size (documents shard with no data-path collective): weak scaling.

`value`   device-resident input -> device-resident (start,end) arrays via jb_cut_device, CUDA events
          on the launching stream, max over ranks.
`e2e`     the same batch through the public host-memory call with HOST buffers, everything inside the
          timed region: pinned text H2D, pipeline, result D2H, and a read of the result on the host.
          The call is jb_cut_batch_bits (result = token start/end bitmaps + per-document token offsets:
          2 bits per input byte come back instead of 8 bytes per token).  `e2e_arrays` is the same through
          jb_cut_batch ((start,end) uint32 arrays), `e2e_pageable` is jb_cut_batch_bits on PAGEABLE text
          (what a Go string is): the library stages it through pinned buffers with host threads.
`config5` (every N) BASELINE config 5 as stated: ONE deterministic 16e9-byte OOV corpus (16 segments of 1e9
          bytes, seed = segment index), HMM on, documents sharded over the N ranks (16/N segments each), device-
          resident rate + e2e + a per-rank oracle parity sample + token count / checksum reduced over ranks
          (identical for every N).
`roofline` algorithmic bytes A = B_in + 8*T_out / duration of the dominant kernel
          (per-kernel CUDA events recorded by the library, jb_profile_*), against the measured HBM
          copy peak in MEASURED_PEAKS.json.  `pipeline_frac` is A / all kernels of the step.
`cpu_baseline` the C restatement of jieba-go (oracle/, kind "port": no Go toolchain here) on all host
          cores over a bounded sample of the same batch; the same sample is a parity check.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    2: dict(kind="freq", hmm=False, name="config2: synthetic corpus sampled from dict frequencies, HMM off"),
    3: dict(kind="oov", hmm=True, name="config3: OOV-rich synthetic corpus (30% random Han), HMM on"),
    4: dict(kind="long", hmm=True, name="config4: 10k-rune unpunctuated blocks, HMM on"),
    5: dict(kind="oov", hmm=True, name="config5: OOV-rich corpus doc-sharded over GPUs, HMM on"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons DURING the timed region (NVML, every 5 ms)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self._nv = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self._nv = nv
            self._h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
            self._names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
        except Exception as e:  # pragma: no cover
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def _sample(self):
        nv = self._nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for bit, name in self._names.items():
            if r & bit:
                self.reasons.add(name)

    def run(self):
        if self._nv is None:
            return
        try:
            while not self._halt.is_set():
                self._sample()
                time.sleep(0.005)
        except Exception as e:  # pragma: no cover
            self.reasons.add("nvml_error:%s" % type(e).__name__)

    def stop(self):
        if self._nv is not None and self.is_alive():
            try:
                self._sample()  # at least one sample taken while the GPU is still busy/just finished
            except Exception:
                pass
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def build_oracle(sd, emit):
    from oracle import c_oracle as co
    pd = co.Dict.from_lines(sd.dict_txt(), 1)
    hm = co.Hmm()
    st = np.array(["BMES".index(s) for s in "BMES" for _ in emit[s]], dtype=np.uint8)
    ru = np.array([c for s in "BMES" for c in emit[s]], dtype=np.uint32)
    va = np.array([v for s in "BMES" for v in emit[s].values()], dtype=np.float64)
    hm.set_emit_arrays(st, ru, va)
    return co, co.Tokenizer(pd, hm)


def cpu_sample(text_np, off_np, target_bytes):
    """Leading whole documents up to ~target_bytes."""
    d = int(np.searchsorted(off_np, target_bytes, side="right")) - 1
    d = max(1, min(d, len(off_np) - 1))
    return text_np[: int(off_np[d])], off_np[: d + 1]


def run_cpu_baseline(otk, co, text_np, off_np, hmm, seconds=12.0):
    cores = co.num_procs()
    probe_t, probe_o = cpu_sample(text_np, off_np, min(len(text_np), 2_000_000 * max(1, cores // 8)))
    t0 = time.perf_counter()
    otk.cut_batch(probe_t, probe_o, hmm, cores)
    dt = max(1e-4, time.perf_counter() - t0)
    rate = len(probe_t) / dt
    st, so = cpu_sample(text_np, off_np, min(len(text_np), int(rate * seconds)))
    t0 = time.perf_counter()
    res = otk.cut_batch(st, so, hmm, cores)
    dt = time.perf_counter() - t0
    return len(st) / dt / 1e6, cores, st, so, res


def run_config5(args, tk, L, sd, emit, dev, rank, world, barrier, stream):
    """BASELINE config 5: one deterministic corpus of 16 segments (seed = segment index, so it is the same corpus for
    every N), HMM on, documents sharded contiguously: rank r owns segments [16 r / N, 16 (r + 1) / N).  Each segment
    is one < 2 GiB device call.  Returns the dict for the bench line on rank 0 (None elsewhere)."""
    import torch
    from jieba_go_b200 import synth
    from jieba_go_b200.dist import reduce_max_sum
    nseg = 16
    seg_bytes = args.config5_bytes // nseg
    mine = list(range(nseg * rank // world, nseg * (rank + 1) // world))
    segs = [synth.make_corpus(sd, "oov", seg_bytes, synth.SEED_BASE + 5000 + i, device=dev) for i in mine]
    cap = max(t.numel() for t, _ in segs) // 3 + 4096
    d_start = torch.empty(cap, dtype=torch.int32, device=dev)
    d_end = torch.empty(cap, dtype=torch.int32, device=dev)
    d_dto = torch.empty(max(o.numel() for _, o in segs), dtype=torch.int64, device=dev)
    d_nt = torch.zeros(2, dtype=torch.int64, device=dev)
    steps, warm = max(1, min(args.steps, 3)), 1
    n_tok, chk = 0, 0

    def one_pass(check):
        nonlocal n_tok, chk
        n_tok, chk = 0, 0
        for t, o in segs:
            tk.cut_device(t, o, True, d_start, d_end, d_dto, d_nt, stream=stream)
            if check:
                stream.synchronize()
                n, status = d_nt.tolist()
                assert status == 0 and n <= cap
                n_tok += n
                # checksum of the segment's (start, end) pairs, reduced mod a prime PER SEGMENT: the sum over segments
                # is then the same number for every sharding of the corpus (and exact in float64)
                chk += int(((d_start[:n].long() * 1_000_003 + d_end[:n].long()) % 2_147_483_629).sum().item()) % 2_147_483_647

    with torch.cuda.stream(stream):
        for _ in range(warm):
            one_pass(False)
    stream.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(steps):
            one_pass(False)
        ev1.record(stream)
    stream.synchronize()
    barrier()
    t_ms = ev0.elapsed_time(ev1) / steps
    with torch.cuda.stream(stream):
        one_pass(True)
    my_bytes = sum(t.numel() for t, _ in segs)
    # e2e: every segment through jb_cut_batch_bits from ONE pinned host buffer (filled outside the timed region)
    e_dt = 0.0
    if not args.no_e2e:
        h_buf = torch.empty(max(t.numel() for t, _ in segs), dtype=torch.uint8).pin_memory()
        for it in range(2):
            e_dt = 0.0
            for t, o in segs:
                h_buf[: t.numel()].copy_(t)
                torch.cuda.synchronize(dev)
                h_off = o.cpu().numpy().astype(np.uint64)
                t0 = time.perf_counter()
                with tk.cut_batch_bits(h_buf[: t.numel()].numpy(), h_off, True) as r:
                    nt = r.n_tokens + int(r.doc_tok_off[-1]) * 0
                e_dt += time.perf_counter() - t0
        del h_buf
    # parity sample on every rank: the first ~24 MB of its first segment against the oracle
    same = None
    if not args.no_cpu:
        co, otk = build_oracle(sd, emit)
        t_np = segs[0][0][:30_000_000].cpu().numpy()
        o_np = segs[0][1].cpu().numpy().astype(np.uint64)
        st, so = cpu_sample(t_np, o_np, 24_000_000)
        res = otk.cut_batch(st, so, True, co.num_procs())
        g = tk.cut_batch(st, so, True)
        same = bool(np.array_equal(g[0], res[0]) and np.array_equal(g[1], res[1]) and np.array_equal(g[2], res[3]))
    (t_max, e_max), (tot_bytes, tot_tok, tot_chk, n_same) = reduce_max_sum(
        [t_ms, e_dt], [float(my_bytes), float(n_tok), float(chk), float(1 if same else 0)], device=dev)
    del segs
    if rank != 0:
        return None
    return {"workload": "config5: ONE %d-byte OOV corpus (16 seeded segments), HMM on, documents sharded over %d GPU(s), %d segment(s) per GPU, one < 2 GiB device call per segment" % (
                int(tot_bytes), world, nseg // world),
            "value": tot_bytes / (t_max * 1e-3) / 1e6, "unit": "MB/s", "ms_per_pass": t_max,
            "e2e": (tot_bytes / e_max / 1e6) if e_max else None, "e2e_ms_per_pass": e_max * 1e3 if e_max else None,
            "bytes": int(tot_bytes), "tokens": int(tot_tok), "checksum": int(tot_chk),
            "parity": None if same is None else ("bit-exact on every rank's 24 MB sample" if int(n_same) == world else "MISMATCH on %d rank(s)" % (world - int(n_same))),
            "steps": steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--bytes", type=int, default=1_000_000_000, help="bytes per GPU per step")
    ap.add_argument("--dict-words", type=int, default=349_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--config5-bytes", type=int, default=16_000_000_000, help="total bytes of the config-5 corpus (16 segments)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON: whatever libraries print there (NCCL's version banner, ...) is sent to
    # stderr by pointing fd 1 at fd 2 for the run; the JSON line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit_line(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import torch
    from jieba_go_b200 import synth

    if args.impl == "reference":
        # the reference's own CPU implementation of the path.  The Go toolchain is absent in this image,
        # so this is the oracle port (C restatement of tokenizer.go), all host threads, bounded sample.
        if rank != 0:
            return
        sd = synth.make_dictionary(n_words=args.dict_words, seed=synth.SEED_BASE)
        emit = synth.make_emit(sd)
        co, otk = build_oracle(sd, emit)
        cores = co.num_procs()
        # the same workload as the GPU arm: the whole --bytes batch every step (about 5 s of all-core work per GB)
        text, doc_off = synth.make_corpus(sd, cfg["kind"], args.bytes, synth.SEED_BASE + args.config, device="cpu")
        t = text.numpy()
        off = doc_off.numpy().astype(np.uint64)
        res = None
        for _ in range(args.warmup):
            res = otk.cut_batch(t, off, cfg["hmm"], cores)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = otk.cut_batch(t, off, cfg["hmm"], cores)
        dt = (time.perf_counter() - t0) / max(1, args.steps)
        v = t.size / dt / 1e6
        n_tok_ref = int(len(res[0])) if res is not None else None
        line = {
            "impl": "reference", "metric": "UTF-8 MB/s segmented", "value": v, "unit": "MB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "bytes_per_gpu_per_step": int(t.size), "docs_per_gpu": int(off.size - 1),
                       "tokens_per_gpu": n_tok_ref, "hmm": cfg["hmm"], "dict_words": args.dict_words,
                       "l2": "n/a (CPU)", "sharding": "documents over host threads"},
            "cpu_baseline": {"value": v, "unit": "MB/s", "cores": cores, "kind": "port",
                             "sample": "the whole batch (%d B) per step, C restatement of jieba-go (not Go; no Go toolchain in the image), %d threads" % (t.size, cores)},
            "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit_line(line)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from jieba_go_b200 import _capi
    from jieba_go_b200.tokenizer import Tokenizer

    sd = synth.make_dictionary(n_words=args.dict_words, seed=synth.SEED_BASE)
    emit = synth.make_emit(sd)
    tk = Tokenizer.from_dict_text(sd.dict_txt(), 1, emit, device=local_rank)  # host batches: default 128 MiB sub-batches, three in flight
    L = _capi.lib()
    # each rank cuts its own shard (documents are independent; no collective on the data path)
    text, doc_off = synth.make_corpus(sd, cfg["kind"], args.bytes, synth.SEED_BASE + args.config + 1000 * rank, device=dev)
    nbytes = text.numel()
    ndocs = doc_off.numel() - 1
    cap = nbytes // 3 + 4096
    d_start = torch.empty(cap, dtype=torch.int32, device=dev)
    d_end = torch.empty(cap, dtype=torch.int32, device=dev)
    d_dto = torch.empty(ndocs + 1, dtype=torch.int64, device=dev)
    d_nt = torch.zeros(2, dtype=torch.int64, device=dev)
    stream = torch.cuda.Stream(device=dev)
    hmm = cfg["hmm"]

    def step():
        tk.cut_device(text, doc_off, hmm, d_start, d_end, d_dto, d_nt, stream=stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
    stream.synchronize()
    n_tok, status = d_nt.tolist()
    assert status == 0, "pipeline status %d" % status
    assert n_tok <= cap, "token capacity too small"

    import ctypes as C
    L.jb_profile_enable(tk.handle, 1)
    ms = (C.c_double * L.jb_profile_num_kernels())()
    nsteps = C.c_uint64()
    L.jb_profile_read(tk.handle, ms, C.byref(nsteps), 1)
    sampler = ClockSampler(local_rank)
    launches0 = L.jb_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
    stream.synchronize()
    barrier()
    clocks = sampler.stop()
    launches = L.jb_kernel_launch_count() - launches0
    t_ms = ev0.elapsed_time(ev1)
    L.jb_profile_read(tk.handle, ms, C.byref(nsteps), 1)
    L.jb_profile_enable(tk.handle, 0)
    kern_ms = {L.jb_profile_kernel_name(i).decode(): ms[i] / max(1, nsteps.value) for i in range(len(ms))}

    # ---- e2e: host buffers in, result on the host, copies inside the timed region ----------------------
    L.jb_bind_thread_to_device(local_rank)  # pinned buffers and the driving thread on the GPU's NUMA node (best effort)
    e2e = None
    e2e_extra = {}
    if not args.no_e2e:
        h_text = text.cpu().pin_memory()
        h_off = doc_off.cpu().numpy().astype(np.uint64)
        h_np = h_text.numpy()
        e_steps = max(2, min(args.steps, 3))

        def time_e2e(call, warm=2):
            for _ in range(warm):  # warm-up (workspaces of the pipeline slots + pinned result buffers)
                call()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                chk = call()
            torch.cuda.synchronize(dev)
            return (time.perf_counter() - t0) / e_steps, chk

        def call_bits(src):
            # the call a user makes: host text in, token bitmaps + per-document offsets out (zero-copy views of the
            # library's pinned result, as the Go shim walks them); the result is READ here: the last document's
            # tokens are expanded from the bits
            with tk.cut_batch_bits(src, h_off, hmm) as r:
                lo = int(h_off[-2] - h_off[0])
                last = int(np.unpackbits(r.start_bits[lo // 32:].view(np.uint8), bitorder="little").sum()) if r.n_bytes else 0
                return r.n_tokens, int(r.doc_tok_off[-1]), last

        def call_arrays():
            with tk.cut_batch_view(h_np, h_off, hmm) as r:
                return r.n_tokens, int(r.end[-1]) if r.n_tokens else 0

        e_dt, chk = time_e2e(lambda: call_bits(h_np))
        assert chk[0] == n_tok and chk[1] == n_tok
        e2e = (e_dt, nbytes + 8 * (ndocs + 1), 2 * 4 * ((nbytes + 31) // 32) + 8 * (ndocs + 1))
        a_dt, chk = time_e2e(call_arrays, warm=2)
        assert chk[0] == n_tok
        e2e_extra["e2e_arrays"] = (a_dt, nbytes + 8 * (ndocs + 1), 8 * n_tok + 8 * (ndocs + 1))
        pageable = np.array(h_np, copy=True)  # ordinary (pageable) host memory, like a Go string
        p_dt, chk = time_e2e(lambda: call_bits(pageable), warm=2)
        assert chk[0] == n_tok
        e2e_extra["e2e_pageable"] = (p_dt, nbytes + 8 * (ndocs + 1), e2e[2])
        del h_text, pageable

    # ---- config 5 as stated: one 16e9-byte corpus, documents sharded over the ranks -----------------------
    c5 = None
    if not args.no_config5 and world in (1, 2, 4, 8, 16):
        c5 = run_config5(args, tk, L, sd, emit, dev, rank, world, barrier, stream)

    # ---- reductions over ranks (max time, total bytes) ---------------------------------------------
    from jieba_go_b200.dist import reduce_max_sum
    ex = [e2e_extra[k][0] if k in e2e_extra else 0.0 for k in ("e2e_arrays", "e2e_pageable")]
    (t_ms_max, e_dt_max, a_dt_max, p_dt_max), (tot_bytes, tot_tok) = reduce_max_sum(
        [t_ms, e2e[0] if e2e else 0.0] + ex, [float(nbytes), float(n_tok)], device=dev)

    if rank == 0:
        peak, peak_src = peaks()
        ms_step = t_ms_max / args.steps
        value = tot_bytes / (ms_step * 1e-3) / 1e6
        A = nbytes + 8 * n_tok  # algorithmic bytes of one launch on this rank
        dom = max(kern_ms, key=lambda k: kern_ms[k])
        t_k = sum(kern_ms.values())
        # DRAM traffic of the dominant kernel per launch: from the committed `ncu --set full` capture of the same
        # workload (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), else null
        traffic, capture = None, {}
        try:
            for rec in json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))):
                if dom.startswith(rec["kernel"]) and rec["config"] == args.config and abs(rec["input_bytes"] - nbytes) <= 0.01 * nbytes:
                    traffic = rec["dram_read_bytes"] + rec["dram_write_bytes"]
                    capture = {k: rec[k] for k in ("capture", "l2_hit_rate_pct", "l1_hit_rate_pct", "lts_throughput_pct") if k in rec}
        except Exception:
            pass
        roof = {
            "bound": "hbm", "kernel": dom, "achieved": A / (kern_ms[dom] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": A / (kern_ms[dom] * 1e-3) / 1e9 / peak, "traffic": traffic,
            "traffic_source": "committed ncu --set full capture of this workload (not measured in this run)" if traffic else None,
            "ncu": capture, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": A,
            "pipeline_achieved": A / (t_k * 1e-3) / 1e9, "pipeline_frac": A / (t_k * 1e-3) / 1e9 / peak,
            "kernel_ms": {k: round(v, 4) for k, v in kern_ms.items()},
        }
        line = {
            "metric": "UTF-8 MB/s segmented", "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "bytes_per_gpu_per_step": nbytes, "docs_per_gpu": ndocs, "tokens_per_gpu": n_tok,
                       "hmm": hmm, "dict_words": args.dict_words, "l2": "inputs (%.0f MB) larger than L2, no flush" % (nbytes / 1e6),
                       "sharding": "documents, one shard per GPU, no collective"},
            "roofline": roof, "clocks": clocks, "gpu_launches": int(launches),
        }
        if e2e:
            line["e2e"] = {"value": tot_bytes / e_dt_max / 1e6, "unit": "MB/s", "h2d_bytes_per_step": int(e2e[1]),
                           "d2h_bytes_per_step": int(e2e[2]), "ms_per_step": e_dt_max * 1e3,
                           "call": "jb_cut_batch_bits: pinned host text in, token start/end bitmaps + doc_tok_off out"}
            for k, dtm, what in (("e2e_arrays", a_dt_max, "jb_cut_batch: (start,end) uint32 arrays out"),
                                 ("e2e_pageable", p_dt_max, "jb_cut_batch_bits on pageable host text (staged through pinned buffers)")):
                if k in e2e_extra:
                    line[k] = {"value": tot_bytes / dtm / 1e6, "unit": "MB/s", "h2d_bytes_per_step": int(e2e_extra[k][1]),
                               "d2h_bytes_per_step": int(e2e_extra[k][2]), "ms_per_step": dtm * 1e3, "call": what}
        if c5:
            line["config5"] = c5
        if not args.no_cpu:
            co, otk = build_oracle(sd, emit)
            t_np = text.cpu().numpy()
            o_np = doc_off.cpu().numpy().astype(np.uint64)
            mbps, cores, st, so, res = run_cpu_baseline(otk, co, t_np, o_np, hmm)
            # the same sample is the parity check
            g = tk.cut_batch(st, so, hmm)
            same = np.array_equal(g[0], res[0]) and np.array_equal(g[1], res[1]) and np.array_equal(g[2], res[3])
            line["cpu_baseline"] = {"value": mbps, "unit": "MB/s", "cores": cores, "kind": "port",
                                    "sample": "first %d B (%d docs) of the batch, C restatement of jieba-go (not Go), %d threads" % (
                                        len(st), len(so) - 1, cores)}
            line["parity"] = "bit-exact on the cpu_baseline sample" if same else "MISMATCH on the cpu_baseline sample"
        emit_line(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
