"""Multi-GPU parity worker (run under torchrun, one rank per GPU, NCCL):  column-sharded fits through the C ABI must give
the same answer as the single-GPU call on the whole design, on every rank.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_worker.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FAM = {"gaussian": (1, 1), "binomial": (2, 2), "poisson": (3, 2), "cox": (4, 3)}
# family, n, p, k, path_type, is_cv, K, ic_type, smax, screening, always
CASES = [
    ("gaussian", 300, 4001, 6, 1, True, 3, 1, 8, 0, ()),
    ("gaussian", 300, 4001, 6, 1, False, 3, 3, 8, 200, ()),
    ("gaussian", 200, 1500, 5, 2, True, 4, 1, 10, 0, (7, 1203)),
    ("binomial", 400, 3000, 5, 2, True, 3, 1, 9, 0, ()),
    ("binomial", 300, 2500, 4, 1, False, 3, 3, 6, 120, (2400,)),
    ("poisson", 400, 2002, 5, 1, True, 2, 1, 6, 0, ()),
    ("cox", 301, 2000, 5, 1, True, 2, 1, 6, 0, ()),
    ("cox", 250, 1800, 4, 1, False, 2, 2, 5, 90, ()),
    ("gaussian", 120, 37, 4, 1, False, 2, 3, 12, 0, ()),  # shards narrower than the support size
]


def main():
    import torch
    import torch.distributed as dist
    from bess_b200 import cbess
    from bess_b200 import dist as bdist
    from bess_b200.gen_data import gen_data
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    failures = []
    for ci, (fam, n, p, k, path_type, is_cv, K, ic_type, smax, scr, always) in enumerate(CASES):
        model_type, data_type = FAM[fam]
        d = gen_data(n, p, fam, k, seed=100 + ci)
        w = np.random.default_rng(ci).uniform(0.5, 1.5, n) if ci % 2 else np.ones(n)
        seq = np.arange(max(1, len(always)), smax + 1)
        s_min = int(seq.min())
        fold = cbess.cv_fold_ids(n, K, 7) if is_cv else None
        ref = cbess.fit(d.x, d.y, data_type, w, True, 1, model_type, 20, 2, path_type, True, ic_type, is_cv, K, seq, s_min,
                        smax, scr > 0, max(scr, 1), always_select=always, fold_of_row=fold, device=local, want_trace=False)
        lo, hi = bdist.shard_range(p, world, rank)
        xs = np.ascontiguousarray(d.x[:, lo:hi])
        out = bdist.fit_column_sharded(xs, lo, p, d.y, w, data_type, True, model_type, 20, path_type, True, ic_type, is_cv,
                                       K, seq, s_min, smax, scr, fold_of_row=fold, device=local, always_select=always)
        sa, sb = np.nonzero(out["beta"])[0], np.nonzero(ref["beta"])[0]
        scale = max(np.abs(ref["beta"]).max(), 1e-300)
        err = float(np.abs(out["beta"] - ref["beta"]).max() / scale)
        ok = (sa.tolist() == sb.tolist() and err < 1e-9 and out["s"] == ref["s"]
              and abs(out["ic"] - ref["ic"]) <= 1e-9 * max(1.0, abs(ref["ic"]))
              and abs(out["coef0"] - ref["coef0"]) <= 1e-9 * max(1.0, abs(ref["coef0"]))
              and out["stats"]["n_pdas_iters"] == ref["stats"]["n_pdas_iters"]
              and out["stats"]["n_boundary_ties"] == 0)
        if scr > 0:
            ok = ok and out["screening_A"].tolist() == ref["screening_A"].tolist()
        # every rank must hold the bit-identical answer
        t = torch.from_numpy(out["beta"]).cuda()
        g = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        same = all(torch.equal(g[0], gi) for gi in g)
        if rank == 0:
            print(f"case {ci} {fam} n={n} p={p} path={path_type} cv={is_cv} scr={scr}: support={sa.tolist()} "
                  f"err={err:.2e} ranks_identical={same} {'OK' if ok and same else 'FAIL'}", flush=True)
        if not (ok and same):
            failures.append(ci)
    # ---- repeated K-fold CV over the ranks (ext.cv_reduce_over_ranks): every rank draws its own folds, the library averages
    # the per-level CV losses with one all-reduce and all ranks must return the same model -- the one a single GPU picks
    # from the average of the per-repetition CV curves
    n, p, K, smax, scr = 300, 4000, 3, 8, 200
    d = gen_data(n, p, "gaussian", 6, seed=77)
    w = np.ones(n)
    seq = np.arange(1, smax + 1)
    lo, hi = bdist.shard_range(p, world, rank)
    xs = np.ascontiguousarray(d.x[:, lo:hi])
    out = bdist.fit_column_sharded(xs, lo, p, d.y, w, 1, True, 1, 20, 1, True, 1, True, K, seq, 1, smax, scr,
                                   cv_seed=500 + rank, device=local, cv_reduce_over_ranks=True)
    curves, models = [], []
    for r in range(world):  # the same repetitions one after another on this GPU, whole design
        o = cbess.fit(d.x, d.y, 1, w, True, 1, 1, 20, 2, 1, True, 1, True, K, seq, 1, smax, True, scr, cv_seed=500 + r,
                      device=local, want_trace=True)
        curves.append(o["ic_all"])
        models.append(o)
    mean = np.mean(np.array(curves), axis=0)
    best = int(np.argmin(mean))
    exp_beta = models[0]["beta_all"][best]  # the full-data chain does not depend on the folds
    scale = max(np.abs(exp_beta).max(), 1e-300)
    ok = (out["s"] == int(seq[best]) and np.nonzero(out["beta"])[0].tolist() == np.nonzero(exp_beta)[0].tolist()
          and float(np.abs(out["beta"] - exp_beta).max() / scale) < 1e-9 and abs(out["ic"] - mean[best]) <= 1e-9 * abs(mean[best]))
    t = torch.from_numpy(out["beta"]).cuda()
    g = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    same = all(torch.equal(g[0], gi) for gi in g)
    if rank == 0:
        print(f"repeated CV over {world} ranks: chosen s={out['s']} expected {int(seq[best])} ranks_identical={same} "
              f"{'OK' if ok and same else 'FAIL'}", flush=True)
    if not (ok and same):
        failures.append("repeated_cv")
    # ---- fold-sharded calls (ext.fold_shard, SURVEY 8e axis A inside ONE call): whole design on every rank, the K fold
    # chains dealt over the ranks, only the fold losses all-reduced -- bit-identical to the single-GPU call on every rank
    FOLD_CASES = [("gaussian", 300, 900, 5, 1, 5, 8, 0), ("gaussian", 260, 3000, 5, 1, 4, 8, 150), ("gaussian", 200, 1500, 5, 2, 5, 10, 0),
                  ("binomial", 400, 800, 4, 1, 5, 6, 0), ("binomial", 360, 700, 4, 2, 3, 9, 0), ("poisson", 400, 600, 4, 1, 4, 5, 0),
                  ("cox", 300, 500, 4, 2, 3, 6, 0),
                  ("gaussian", 200, 400, 4, 2, 2, 6, 0)]  # K = 2: with 4 or 8 ranks some ranks own no fold chain at all
    for fi, (fam, n, p, k, path_type, K, smax, scr) in enumerate(FOLD_CASES):
        model_type, data_type = FAM[fam]
        d = gen_data(n, p, fam, k, seed=300 + fi)
        w = np.ones(n)
        seq = np.arange(1, smax + 1)
        fold = cbess.cv_fold_ids(n, K, 11)
        ref = cbess.fit(d.x, d.y, data_type, w, True, 1, model_type, 20, 2, path_type, True, 1, True, K, seq, 1, smax, scr > 0,
                        max(scr, 1), fold_of_row=fold, device=local, want_trace=False)
        out = bdist.fit_fold_sharded(d.x, d.y, w, data_type, True, model_type, 20, path_type, True, 1, K, seq, 1, smax, scr,
                                     fold_of_row=fold, device=local)
        # (bit-identical whenever the chains get the same cluster size as in the single-GPU batch; a rank with fewer chains
        # may give each a wider cluster -- then the Gram sums run over different row slices and the last bits move)
        bit = np.array_equal(out["beta"], ref["beta"]) and out["ic"] == ref["ic"]
        scale = max(np.abs(ref["beta"]).max(), 1e-300)
        ok = (np.nonzero(out["beta"])[0].tolist() == np.nonzero(ref["beta"])[0].tolist() and out["s"] == ref["s"]
              and float(np.abs(out["beta"] - ref["beta"]).max() / scale) < 1e-9
              and abs(out["ic"] - ref["ic"]) <= 1e-9 * max(1.0, abs(ref["ic"]))
              and abs(out["coef0"] - ref["coef0"]) <= 1e-9 * max(1.0, abs(ref["coef0"])))
        fewer = out["stats"]["n_fits"] < ref["stats"]["n_fits"] or K <= 2  # this rank fitted only its share of the fold chains
        t = torch.from_numpy(out["beta"]).cuda()
        g = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        same = all(torch.equal(g[0], gi) for gi in g)
        if rank == 0:
            print(f"fold-sharded {fi} {fam} path={path_type} K={K} scr={scr}: s={out['s']} ic={out['ic']:.6f} fits {out['stats']['n_fits']}"
                  f"/{ref['stats']['n_fits']} bit_identical={bit} ranks_identical={same} {'OK' if ok and same and fewer else 'FAIL'}",
                  flush=True)
        if not (ok and same and fewer):
            failures.append(f"fold{fi}")
    dist.barrier()
    dist.destroy_process_group()
    if failures:
        print(f"rank {rank}: FAILED cases {failures}", flush=True)
        sys.exit(1)
    if rank == 0:
        print("ALL SHARDED CASES OK", flush=True)


if __name__ == "__main__":
    main()
