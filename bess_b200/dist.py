"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL on the GPU box, gloo in CPU tests).

The path shards along two axes:
  A. chains  -- the 1+K warm-start lineages (full fit + CV folds) are independent; rank r owns chains c with
                c % world == r; only the K fold losses per path step are reduced.
  B. columns -- for very large p, rank r owns the contiguous column range ``shard_range(p, world, r)``; every PDAS
                iteration each rank runs the dual sweep + exact local top-k on its columns, the (value, index)
                candidates are all-gathered and merged identically on every rank (``global_topk_from_local``).
On top of both sits the embarrassingly parallel axis the bench scales on: repeated K-fold CV -- rank r runs the CV path with
its own fold assignment (``cv_seed + r``), the per-level CV losses are averaged over ranks (``repeated_cv_reduce``), every
rank picks the same sparsity level from the averaged curve.
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
    t = torch.from_numpy(np.array(x, dtype=np.float64, copy=True)).to(dev)  # never reduce into the caller's array
    dist.all_reduce(t)
    return t.cpu().numpy()


_NCCL_ID = None


def nccl_unique_id() -> bytes:
    """The job's NCCL unique id for the library's own communicator: created on rank 0 through the C ABI
    (``bess_b200_nccl_unique_id``), broadcast once over the existing process group, cached."""
    global _NCCL_ID
    if _NCCL_ID is None:
        dist = _dist()
        obj = [None]
        if dist.get_rank() == 0:
            buf = C.create_string_buffer(128)
            _lib.check(_lib.load().bess_b200_nccl_unique_id(C.cast(buf, C.c_void_p)))
            obj = [bytes(buf.raw)]
        dist.broadcast_object_list(obj, src=0)
        _NCCL_ID = obj[0]
    return _NCCL_ID


def fit_column_sharded(x_shard, col_lo, p_total, y, weight, data_type, is_normal, model_type, max_iter, path_type,
                       is_warm_start, ic_type, is_cv, K, sequence, s_min, s_max, screening_size, cv_seed=123,
                       fold_of_row=None, device=None, x_shard_device_ptr=None, n=None, p_local=None, profile=False,
                       always_select=(), want_curve=False, cv_reduce_over_ranks=False):
    """Multi-GPU fit with the columns of X sharded across the ranks of the current process group (one ``bess_b200_fit``
    call per rank, the library talks NCCL itself).

    Each rank holds columns [col_lo, col_lo + p_local) = ``shard_range(p_total, world, rank)`` (host array ``x_shard`` --
    uploaded over the rank's own PCIe link -- or an HBM-resident shard via ``x_shard_device_ptr``).
      * ``screening_size > 0`` (BASELINE config 5): local marginal-utility sweep + exact local top-k, NCCL all-gather of
        (utility, global index) candidates, identical merge, all-reduce of the kept columns; the PDAS path then runs
        replicated on the n x screening_size design.
      * ``screening_size == 0``: every PDAS iteration is sharded -- local dual sweep + local top-k, all-gather of the
        candidates, all-reduce of the k active columns, replicated active-set fit.
    Returns the same dict as ``cbess.fit`` (beta has length p_total), identical on every rank -- unless the ranks pass
    different ``cv_seed`` / ``fold_of_row`` with ``screening_size > 0``: the screening is then still joint, and each rank
    runs its own CV repetition on the replicated screened design.  With ``cv_reduce_over_ranks=True`` the library then
    averages the per-level CV losses over the ranks itself (one NCCL all-reduce) before choosing the level, so the ranks
    again return the same model: repeated K-fold CV, one repetition per GPU (``repeated_cv_reduce`` is the same
    reduction done by the caller with ``torch.distributed``)."""
    import torch
    from . import cbess
    dist = _dist()
    if device is None:
        device = torch.cuda.current_device()
    scr = int(screening_size or 0)
    return cbess.fit(x_shard, y, data_type, weight, is_normal, 1, model_type, max_iter, 2, path_type, is_warm_start,
                     ic_type, is_cv, K, sequence, s_min, s_max, scr > 0, max(scr, 1), always_select=always_select,
                     fold_of_row=fold_of_row, cv_seed=cv_seed, device=device, x_device_ptr=x_shard_device_ptr, n=n,
                     p=p_local, want_trace=False, profile=profile, world=dist.get_world_size(), rank=dist.get_rank(),
                     col_lo=col_lo, p_total=p_total, nccl_id=nccl_unique_id(), want_curve=want_curve,
                     cv_reduce_over_ranks=cv_reduce_over_ranks)


def fold_shard_chains(K, world, rank, path_type=1):
    """The chains rank ``rank`` runs in a fold-sharded call and which fold losses it contributes (host helper, no GPU)."""
    ch = np.zeros(K + 1, dtype=np.int32)
    cnt = np.zeros(K + 1, dtype=np.int32)
    n = _lib.load().bess_b200_fold_shard_chains(int(K), int(world), int(rank), int(path_type),
                                                ch.ctypes.data_as(_lib.ip), cnt.ctypes.data_as(_lib.ip))
    if n < 0:
        raise ValueError(_lib.last_error())
    return ch[:n].tolist(), cnt[:n].tolist()


def fit_fold_sharded(x, y, weight, data_type, is_normal, model_type, max_iter, path_type, is_warm_start, ic_type, K,
                     sequence, s_min, s_max, screening_size=0, algorithm_type=1, cv_seed=123, fold_of_row=None, device=None,
                     x_device_ptr=None, n=None, p=None, profile=False, always_select=(), g_index=None, lambda_seq=(0.0,),
                     want_curve=False):
    """ONE K-fold-CV call spread over the ranks of the current process group by FOLDS (SURVEY 8e axis A): every rank
    holds the whole design and makes this same call; the library deals the K fold chains of ``Metric::test_loss``
    (Metric.h:150-195) over the ranks (chain c on rank c % world), runs the full-data chain everywhere and all-reduces only
    the per-fold test losses (``ext.fold_shard``).  Every rank returns the same model, bit-identical to the single-GPU
    call.  Sequential and golden-section paths, all four families, groups and screening allowed (the screening is then
    done redundantly by every rank on its own copy)."""
    import torch
    from . import cbess
    dist = _dist()
    if device is None:
        device = torch.cuda.current_device()
    scr = int(screening_size or 0)
    return cbess.fit(x, y, data_type, weight, is_normal, algorithm_type, model_type, max_iter, 2, path_type, is_warm_start,
                     ic_type, True, K, sequence, s_min, s_max, scr > 0, max(scr, 1), always_select=always_select,
                     fold_of_row=fold_of_row, cv_seed=cv_seed, device=device, x_device_ptr=x_device_ptr, n=n, p=p,
                     want_trace=False, profile=profile, world=dist.get_world_size(), rank=dist.get_rank(),
                     nccl_id=nccl_unique_id(), want_curve=want_curve, g_index=g_index, lambda_seq=lambda_seq, fold_shard=True)


def repeated_cv_reduce(cv_curve: np.ndarray):
    """Repeated K-fold CV over the ranks of the current process group: ``cv_curve[i]`` is this rank's mean fold loss at
    the i-th sparsity level (its own fold assignment); returns (curve averaged over the ranks, index of its first
    minimum) -- identical on every rank, so all ranks choose the same level.  One small all-reduce per path."""
    dist = _dist()
    mean = allreduce_sum(np.asarray(cv_curve, dtype=np.float64)) / dist.get_world_size()
    return mean, int(np.argmin(mean))


def unique_fits_repeated_cv(world: int, n_levels: int, K: int) -> int:
    """PDAS fits of a ``world``-repetition repeated-CV job that are not duplicates of one another: the full-data chain
    is the same on every rank (counted once), the K fold chains differ per repetition."""
    return n_levels * (1 + world * K)
