"""Host<->device copy floor of the box at N GPUs (what bounds the end-to-end number of bench.py).

    python tools/pcie_probe.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py

Every rank moves what one bench step moves per GB of text -- 1.00 GB host->device, and device->host either the
bitmap result (0.25 GB) or the (start,end) arrays (1.08 GB) -- all ranks at the same time, pinned memory, with and
without binding the rank to its GPU's NUMA node (jb_bind_thread_to_device).  Rank 0 prints one JSON line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ.get("RANK", "0"))
    lrank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(lrank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    from jieba_go_b200 import _capi
    L = _capi.lib()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def node_of_gpu():
        try:
            bus = torch.cuda.get_device_properties(lrank).pci_bus_id
            dom = torch.cuda.get_device_properties(lrank).pci_domain_id
            dev = torch.cuda.get_device_properties(lrank).pci_device_id
            p = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
            return int(open(p).read())
        except Exception:
            return None

    out = {"world": world, "cpus_allowed": len(os.sched_getaffinity(0)), "gpu_numa_node": node_of_gpu()}
    try:
        out["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
    except Exception:
        out["numa_nodes"] = None
    n_in, n_bits, n_arr = 1_000_000_000, 250_000_000, 1_080_000_000
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for bind in (False, True):
        if bind:
            out["bound_to_node"] = L.jb_bind_thread_to_device(lrank)
            out["cpus_after_bind"] = len(os.sched_getaffinity(0))
        h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
        h_in.fill_(1)
        h_out = torch.empty(n_arr, dtype=torch.uint8).pin_memory()
        h_out.fill_(1)
        d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
        d_out = torch.empty(n_arr, dtype=torch.uint8, device="cuda")

        def run(h2d, d2h_bytes, reps=3):
            def once():
                if h2d:
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if d2h_bytes:
                    with torch.cuda.stream(s2):
                        h_out[:d2h_bytes].copy_(d_out[:d2h_bytes], non_blocking=True)
            once()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                once()
            barrier()
            return (time.perf_counter() - t0) / reps * 1e3

        tag = "bound" if bind else "unbound"
        res[tag] = {"h2d_1.00GB_ms": run(True, 0), "d2h_1.08GB_ms": run(False, n_arr), "d2h_0.25GB_ms": run(False, n_bits),
                    "h2d+d2h_bits_ms": run(True, n_bits), "h2d+d2h_arrays_ms": run(True, n_arr)}
        del h_in, h_out, d_in, d_out
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, (out, res))
    else:
        gathered = [(out, res)]
    if rank == 0:
        line = {"probe": "pcie", "n_gpus": world, "ranks": [g[0] for g in gathered]}
        for tag in ("unbound", "bound"):
            line[tag] = {k: max(g[1][tag][k] for g in gathered) for k in gathered[0][1][tag]}
            line[tag]["e2e_floor_bits_GBps_box"] = world * 1.0 / (line[tag]["h2d+d2h_bits_ms"] * 1e-3)
            line[tag]["e2e_floor_arrays_GBps_box"] = world * 1.0 / (line[tag]["h2d+d2h_arrays_ms"] * 1e-3)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
