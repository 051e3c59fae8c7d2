"""Host-side logic of the runner (no GPU): LR schedule, metrics on token paths, rank sharding, and the N>1 metric /
loss reduction over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from openp5_b200 import runner as R
from oracle import p5_oracle as po


def test_schedule_matches_oracle_and_reference_plan():
    total, warm = R.schedule_plan(batches_per_epoch=2704, epochs=10, warmup_prop=0.05, grad_accum=1)
    assert (total, warm) == (27040, 1352)           # SingleRunner.py:181-183
    for s in [0, 1, 100, 1352, 1353, 20000, 27040]:
        assert R.linear_schedule(s, warm, total) == po.linear_schedule(s, warm, total)


def test_metrics_on_token_paths_match_reference_golden(ref_helpers):
    # map the golden's string predictions onto token paths: equality structure is all that matters
    vocab = {s: i + 10 for i, s in enumerate(sorted(set(ref_helpers["preds"] + ref_helpers["targets"])))}
    seqs = [[0, vocab[p], 1, 0] for p in ref_helpers["preds"]]
    gold = [[vocab[t], 1, 0] for t in ref_helpers["targets"]]
    rel = R.rel_results(seqs, ref_helpers["scores"], gold, 4)
    assert rel == ref_helpers["rel"]
    assert np.allclose(R.metric_sums(rel, ["hit@1", "hit@3", "ndcg@3", "ndcg@4"]), ref_helpers["metrics"])


def test_strip_special():
    assert R.strip_special([0, 300, 301, 1, 0, 0]) == (300, 301)


def test_shard_for_rank_covers_and_balances():
    idx = list(range(11))
    shards = [R.shard_for_rank(idx, r, 4) for r in range(4)]
    assert all(len(s) == 3 for s in shards)                       # every rank emits the same number of samples
    assert set(sum(shards, [])) == set(idx)
    assert shards[0][:3] == [0, 4, 8] and shards[3][:2] == [3, 7]   # list[r::world]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank evaluated a different shard of users
    sums = [3.0, 1.5] if rank == 0 else [1.0, 0.5]
    res, total = R.allreduce_metrics(sums, 10 if rank == 0 else 6, device="cpu")
    # gradient averaging semantics the engine implements with NCCL (mean over ranks), checked with gloo
    g = torch.full((4,), float(rank + 1))
    dist.all_reduce(g, op=dist.ReduceOp.SUM)
    g /= world
    q.put((rank, res, total, g.tolist()))
    dist.destroy_process_group()


def test_world2_gloo_metric_and_grad_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    outs = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, res, total, g in outs:
        assert total == 16
        assert np.allclose(res, [4.0 / 16, 2.0 / 16])
        assert g == [1.5] * 4
