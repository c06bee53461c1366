"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL on the GPU box, gloo in CPU tests).

The path shards along two axes:
  A. chains  -- the 1+K warm-start lineages (full fit + CV folds) are independent; rank r owns chains c with
                c % world == r; only the K fold losses per path step are reduced.
  B. columns -- for very large p, rank r owns the contiguous column range ``shard_range(p, world, r)``; every PDAS
                iteration each rank runs the dual sweep + exact local top-k on its columns, the (value, index)
                candidates are all-gathered and merged identically on every rank (``global_topk_from_local``).
The merge rule is the library's total order (larger value first, lower index first), implemented once in the C ABI
(``bess_b200_merge_candidates``) so host and device paths agree."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import dp, ip


def shard_range(p: int, world: int, rank: int):
    lib = _lib.load()
    lo, hi = C.c_longlong(), C.c_longlong()
    lib.bess_b200_shard_range(int(p), int(world), int(rank), C.byref(lo), C.byref(hi))
    return int(lo.value), int(hi.value)


def chain_owner(chain: int, world: int) -> int:
    return int(_lib.load().bess_b200_chain_owner(int(chain), int(world)))


def merge_candidates(vals: np.ndarray, idx: np.ndarray, k: int) -> np.ndarray:
    lib = _lib.load()
    v = np.ascontiguousarray(vals, dtype=np.float64)
    i = np.ascontiguousarray(idx, dtype=np.int32)
    out = np.zeros(k, dtype=np.int32)
    _lib.check(lib.bess_b200_merge_candidates(v.ctypes.data_as(dp), i.ctypes.data_as(ip), v.size, k,
                                              out.ctypes.data_as(ip)))
    return out


def _dist():
    import torch.distributed as dist
    return dist


def global_topk_from_local(local_vals: np.ndarray, local_idx: np.ndarray, k: int) -> np.ndarray:
    """All-gather every rank's local top-k candidates (value, global column index) and merge to the global top-k
    (ascending indices).  Candidate lists are padded to k entries with -1 values so the collective is fixed-size."""
    import torch
    dist = _dist()
    world = dist.get_world_size()
    v = np.full(k, -1.0)
    i = np.full(k, -1, dtype=np.int32)
    v[:local_vals.size] = local_vals
    i[:local_idx.size] = local_idx
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    tv = torch.from_numpy(v).to(dev)
    ti = torch.from_numpy(i).to(dev)
    gv = [torch.empty_like(tv) for _ in range(world)]
    gi = [torch.empty_like(ti) for _ in range(world)]
    dist.all_gather(gv, tv)
    dist.all_gather(gi, ti)
    av = torch.cat(gv).cpu().numpy()
    ai = torch.cat(gi).cpu().numpy()
    keep = ai >= 0
    return merge_candidates(av[keep], ai[keep], k)


def allreduce_sum(x: np.ndarray) -> np.ndarray:
    import torch
    dist = _dist()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
    dist.all_reduce(t)
    return t.cpu().numpy()


def fit_column_sharded(x_shard, col_lo, p_total, y, weight, data_type, is_normal, model_type, max_iter, path_type,
                       is_warm_start, ic_type, is_cv, K, sequence, s_min, s_max, screening_size, cv_seed=123,
                       fold_of_row=None, device=None, x_shard_device_ptr=None, n=None, p_local=None, profile=False):
    """Multi-GPU fit of a SCREENED problem (BASELINE config 5) with the columns of X sharded across ranks.

    Each rank holds columns [col_lo, col_lo + p_local) (host array ``x_shard`` -- uploaded over the rank's own PCIe
    link -- or an HBM-resident shard via ``x_shard_device_ptr``).  Steps:
      1. local marginal-utility sweep + exact local top-k on the rank's columns      (screening.cpp:40-61, 63-66)
      2. all-gather of the (utility, global index) candidates, identical merge        -> screening_A on every rank
      3. every rank copies the kept columns it owns into a shared n x m buffer, all-reduce(sum) over NVLink
      4. the PDAS path + CV on the n x m screened design, replicated on every rank    (bess.cpp:61-185)
    Returns the same dict as ``cbess.fit`` (beta has length p_total), identical on every rank."""
    import torch
    from . import cbess
    from .engine import GpuEngine
    dist = _dist()
    rank = dist.get_rank()
    if device is None:
        device = torch.cuda.current_device()
    eng = GpuEngine(device)
    if x_shard_device_ptr is None:
        n, p_local = x_shard.shape
    eng.load(x_shard, y, weight, model_type, x_device_ptr=x_shard_device_ptr, n=n, p=p_local)
    vals, idx = eng.screen_local(screening_size)
    A = global_topk_from_local(vals, (idx + col_lo).astype(np.int32), screening_size)
    m = int(A.size)
    ld = (m + 1) & ~1
    Xs = torch.zeros((n, ld), dtype=torch.float64, device=f"cuda:{device}")
    torch.cuda.current_stream().synchronize()
    mine = np.nonzero((A >= col_lo) & (A < col_lo + p_local))[0]
    eng.gather_columns(A[mine] - col_lo, mine, Xs.data_ptr(), ld)
    eng.close()
    dist.all_reduce(Xs)
    torch.cuda.current_stream().synchronize()
    xs = Xs if ld == m else Xs[:, :m].contiguous()
    out = cbess.fit(None, y, data_type, weight, is_normal, 1, model_type, max_iter, 2, path_type, is_warm_start, ic_type,
                    is_cv, K, sequence, s_min, s_max, False, 1, fold_of_row=fold_of_row, cv_seed=cv_seed, device=device,
                    x_device_ptr=xs.data_ptr(), n=n, p=m, want_trace=False, profile=profile)
    beta = np.zeros(p_total)
    beta[A] = out["beta"]
    out["beta"] = beta
    out["screening_A"] = A
    del rank
    return out
