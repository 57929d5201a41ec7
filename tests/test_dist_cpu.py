"""N>1 host logic on CPU: world_size-2 gloo process group (documents shard, no data-path collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jieba_go_b200 import synth
from jieba_go_b200.dist import merge_shards, reduce_max_sum, shard_docs

from helpers import c_oracle_tokenizer


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank rebuilds the same (seeded) workload and cuts only its shard -- with the ORACLE here,
    # since there is no GPU on this box; the sharding/merging/reduction logic is what is under test
    sd = synth.make_dictionary(n_words=3000, seed=synth.SEED_BASE + 11, total_freq=1.0e6, max_len=8)
    emit = synth.make_emit(sd, seed=synth.SEED_BASE + 12)
    tk = c_oracle_tokenizer(sd, emit)
    text, doc_off = synth.make_corpus(sd, "oov", 120_000, synth.SEED_BASE + 90)
    t = text.numpy()
    off = doc_off.numpy().astype(np.uint64)
    lo, hi = shard_docs(off, world)[rank]
    sub_off = off[lo:hi + 1] - off[lo]
    sub_text = t[int(off[lo]):int(off[hi])]
    s, e, _, d = tk.cut_batch(sub_text, sub_off, True, 1)
    times, amounts = reduce_max_sum([1.0 + rank, 5.0 - rank], [float(sub_text.size), float(len(s))])
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), s=s, e=e, d=d, nd=hi - lo, times=times, amounts=amounts)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_doc_sharding(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    # reductions: max of times, sum of amounts, identical on both ranks
    for p in parts:
        assert p["times"].tolist() == [2.0, 5.0]
    assert parts[0]["amounts"].tolist() == parts[1]["amounts"].tolist()
    # merged shards == one-shot cut
    sd = synth.make_dictionary(n_words=3000, seed=synth.SEED_BASE + 11, total_freq=1.0e6, max_len=8)
    emit = synth.make_emit(sd, seed=synth.SEED_BASE + 12)
    tk = c_oracle_tokenizer(sd, emit)
    text, doc_off = synth.make_corpus(sd, "oov", 120_000, synth.SEED_BASE + 90)
    t = text.numpy()
    off = doc_off.numpy().astype(np.uint64)
    s, e, _, d = tk.cut_batch(t, off, True, 2)
    ms, me, md = merge_shards([(p["s"], p["e"], p["d"]) for p in parts], [int(p["nd"]) for p in parts])
    assert np.array_equal(ms, s) and np.array_equal(me, e) and np.array_equal(md, d)
    assert parts[0]["amounts"].tolist() == [float(t.size), float(len(s))]


def test_shard_docs_balances_bytes():
    off = np.array([0, 10, 20, 1000, 1010, 2000, 2000, 2010], dtype=np.int64)
    sh = shard_docs(off, 3)
    assert sh[0][0] == 0 and sh[-1][1] == 7 and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
    assert shard_docs(off, 1) == [(0, 7)]
    assert len(shard_docs(np.array([0, 5]), 4)) == 4
