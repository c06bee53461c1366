"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: column shards, local top-k candidate exchange and merge
(SURVEY 8e axis B), and the fold-loss reduction of chain sharding (axis A)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bess_b200 import dist as bdist
        from oracle import pdas_oracle as orc
        rng = np.random.default_rng(3)
        p, k = 10007, 25
        bd = np.floor(rng.random(p) * 1000)  # ties on purpose
        lo, hi = bdist.shard_range(p, world, rank)
        local = bd[lo:hi]
        # local exact top-k (on the GPU build this is the device top-k over the rank's columns)
        kk = min(k, local.size)
        li = orc.max_k(local, kk)
        A = bdist.global_topk_from_local(local[li], (li + lo).astype(np.int32), k)
        ok_topk = A.tolist() == orc.max_k(bd, k).tolist()
        # chain sharding: each rank owns folds {c : c % world == rank}; losses are summed over ranks
        K = 5
        fold_losses = np.arange(1, K + 1, dtype=np.float64) * 1.5
        mine = np.array([fold_losses[c] if bdist.chain_owner(1 + c, world) == rank else 0.0 for c in range(K)])
        red = bdist.allreduce_sum(mine)
        ok_red = np.allclose(red, fold_losses)
        # fold-sharded call (ext.fold_shard): the chain lists of the ranks cover every fold, every fold loss is contributed
        # exactly once, chain 0 runs everywhere, and under golden section every rank also runs the last fold chain
        ok_fs = True
        for Kf, ptype in ((5, 1), (10, 2), (3, 2), (7, 1)):
            ch, cnt = bdist.fold_shard_chains(Kf, world, rank, ptype)
            contrib = np.zeros(Kf + 1)
            for c, ct in zip(ch, cnt):
                contrib[c] += ct
            tot = bdist.allreduce_sum(contrib)
            ok_fs = ok_fs and ch[0] == 0 and cnt[0] == 0 and tot[0] == 0 and np.all(tot[1:] == 1)
            ok_fs = ok_fs and (ptype == 1 or Kf in ch) and all(c % world == rank or (c == Kf and ptype != 1) for c in ch[1:])
        ok_red = ok_red and ok_fs
        # repeated CV: every rank holds its own CV curve; all ranks must agree on the averaged curve and the chosen level
        curves = np.array([[5.0, 3.0, 2.5, 2.6, 4.0], [5.5, 2.0, 2.9, 2.7, 4.5]])
        mean, best = bdist.repeated_cv_reduce(curves[rank])
        ok_rep = np.allclose(mean, curves.mean(axis=0)) and best == int(np.argmin(curves.mean(axis=0)))
        ok_rep = ok_rep and bdist.unique_fits_repeated_cv(world, 20, 10) == 20 * (1 + 2 * 10)
        q.put((rank, ok_topk, ok_red and ok_rep))
    finally:
        dist.destroy_process_group()


def test_two_rank_candidate_merge_and_loss_reduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert all(ok1 and ok2 for _, ok1, ok2 in res), res
