"""Measurement for the SURVEY 8f rows built after the core path: the bsrr Powell search (pgs_path) and group selection,
at sizes well beyond the golden fixtures.  Every case runs through the C ABI on the GPU and through the reference
(oracle/_ref, one host thread as shipped) on the same inputs in the same process; prints both times, the speed-up and the
parity of the two results (support, chosen lambda, coefficients, criterion).  One JSON line per case on stdout."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bess_b200 import cbess  # noqa: E402
from bess_b200.gen_data import gen_data  # noqa: E402
from oracle import ref  # noqa: E402

FAM = {"gaussian": (1, 1), "binomial": (2, 2), "poisson": (3, 2), "cox": (4, 3)}

# name: family, n, p, k, group size (0 = none), algorithm_type, path_type, is_cv, K, ic_type, s_min, s_max, pgs (lmin, lmax, nl, powell) or None
CASES = {
    "bsrr_lm_pgs_gs_cv": ("gaussian", 1000, 5000, 10, 0, 5, 2, True, 5, 1, 1, 20, (0.01, 100.0, 100, 1)),
    "bsrr_lm_pgs_seq_gic": ("gaussian", 1000, 5000, 10, 0, 5, 2, False, 5, 3, 1, 20, (0.01, 100.0, 10, 2)),
    "bsrr_logit_pgs_gs_gic": ("binomial", 1000, 3000, 8, 0, 5, 2, False, 5, 3, 1, 16, (0.001, 1.0, 100, 1)),
    "group_lm_seq_cv": ("gaussian", 1000, 8000, 12, 4, 2, 1, True, 5, 1, 1, 10, None),
    "group_logit_seq_gic": ("binomial", 1000, 4000, 8, 4, 2, 1, False, 5, 3, 1, 8, None),
    "group_cox_seq_gic": ("cox", 600, 2000, 6, 4, 2, 1, False, 5, 3, 1, 6, None),
}


def main():
    only = set(sys.argv[1:])
    for name, (fam, n, p, k, gs, alg, path_type, is_cv, K, ic_type, s_min, s_max, pgs) in CASES.items():
        if only and name not in only:
            continue
        model_type, data_type = FAM[fam]
        d = gen_data(n, p, fam, k, seed=7)
        w = np.ones(n)
        gi = np.arange(0, p, gs, dtype=np.int32) if gs else None
        seq = np.arange(s_min, s_max + 1, dtype=np.int32)
        kw = dict(lambda_min=pgs[0], lambda_max=pgs[1], n_lambda=pgs[2], powell_path=pgs[3]) if pgs else {}
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            out = cbess.fit(d.x, d.y, data_type, w, True, alg, model_type, 20, 2, path_type, True, ic_type, is_cv, K, seq,
                            s_min, s_max, False, 1, cv_seed=123, g_index=gi, want_trace=False, **kw)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        t0 = time.perf_counter()
        r = ref.bess_lambda(d.x, d.y, data_type, w, True, alg, model_type, 20, path_type, True, ic_type, is_cv, K, seq, s_min,
                            s_max, g_index=gi, **kw)
        t_ref = time.perf_counter() - t0
        sa, sb = np.nonzero(out["beta"])[0].tolist(), np.nonzero(r["beta"])[0].tolist()
        scale = max(float(np.max(np.abs(r["beta"]))), 1e-300)
        line = dict(case=name, family=fam, n=n, p=p, group_size=gs, algorithm_type=alg, path_type=path_type, cv=is_cv,
                    gpu_ms=best * 1e3, ref_cpu_s=t_ref, speedup=t_ref / best, fits=out["stats"]["n_fits"],
                    support_equal=sa == sb, n_selected=len(sb),
                    beta_rel_err=float(np.max(np.abs(out["beta"] - r["beta"])) / scale),
                    ic_rel_err=abs(out["ic"] - r["ic"]) / max(abs(r["ic"]), 1e-300),
                    lambda_gpu=out["lam"], lambda_ref=r["lambda_"], ties=out["stats"]["n_boundary_ties"])
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
