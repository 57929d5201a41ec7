"""Multi-GPU plumbing for the Cut path: documents shard, nothing else is exchanged.

Blocks -- a fortiori documents -- are independent (CutParallel cuts them in any order,
/root/reference/tokenizer.go:100-105), so each rank owns a contiguous range of documents and a replica
of the tables; there is no collective on the data path.  torch.distributed is used only to agree on
timing (max over ranks) and totals (sum over ranks).
"""
import numpy as np
import torch


def shard_docs(doc_off, world: int):
    """Split documents into `world` contiguous ranges balanced by bytes.
    doc_off: int array [ndocs+1] -> list of (first_doc, last_doc_exclusive), one per rank."""
    doc_off = np.asarray(doc_off, dtype=np.int64)
    nd = len(doc_off) - 1
    total = int(doc_off[-1] - doc_off[0])
    cuts = [0]
    for r in range(1, world):
        target = doc_off[0] + total * r // world
        d = int(np.searchsorted(doc_off, target, side="left"))
        cuts.append(min(max(d, cuts[-1]), nd))
    cuts.append(nd)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def merge_shards(results, doc_counts):
    """Concatenate per-rank (start, end, doc_tok_off) results in rank order into one result."""
    starts = np.concatenate([r[0] for r in results]) if results else np.zeros(0, np.uint32)
    ends = np.concatenate([r[1] for r in results]) if results else np.zeros(0, np.uint32)
    dto = [np.zeros(1, np.uint64)]
    base = 0
    for (s, e, d), nd in zip(results, doc_counts):
        assert len(d) == nd + 1
        dto.append(d[1:].astype(np.uint64) + np.uint64(base))
        base += int(d[-1])
    return starts, ends, np.concatenate(dto)


def reduce_max_sum(times, amounts, device=None):
    """all-reduce: element-wise MAX of `times`, SUM of `amounts` (lists of floats); no-op without a process group."""
    import torch.distributed as dist
    t = torch.tensor(list(times), dtype=torch.float64, device=device)
    a = torch.tensor(list(amounts), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
    return t.tolist(), a.tolist()
