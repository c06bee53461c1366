"""Generates tests/golden/pgs/*.npz from the REAL reference (oracle/_ref/libbess_ref.so): the Powell search of
pgs_path (/root/reference/src/path.cpp:1138-1309) over (sparsity level, lambda), with golden-section (powell_path 1)
or grid-walk (powell_path 2) line searches.  Run in the build container only:  python tests/golden/pgs/make_pgs.py
Each file holds the inputs, the CV folds the reference drew, bessCpp's outputs incl. the chosen lambda, and the log of
every Algorithm::fit the reference made (sparsity level, lambda, #train rows, coef0_init, nnz(beta_init)) from
oracle/ref_probe.cpp:ref_pgs_trace -- the order of evaluations is part of the contract, the search is stateful."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
from bess_b200.gen_data import gen_data  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FAM = {"gaussian": (1, 1), "binomial": (2, 2), "poisson": (3, 2), "cox": (4, 3)}

# name: family, n, p, k, powell_path, is_cv, K, ic_type, s_min, s_max, lambda_min, lambda_max, n_lambda, warm, weighted, seed
# Cox: lambda_max is kept small.  The reference adds the ridge term to its NEGATIVE-definite Hessian (Algorithm.h:1472),
# so once 2*lambda reaches the smallest eigenvalue of the information matrix the Newton system is indefinite, the
# iteration wanders for all 30 steps and the result is chaotic in the last bits of every intermediate (observed: LU vs
# the reference's LDLT differ by 6e-7 at lambda = 1 on these data) -- no 1e-8 parity is definable there.
CASES = {
    "lm_gs_gic": ("gaussian", 150, 200, 5, 1, False, 5, 3, 1, 12, 0.01, 100.0, 100, True, False, 61),
    "lm_gs_cv": ("gaussian", 150, 200, 5, 1, True, 3, 1, 1, 12, 0.001, 10.0, 100, True, True, 62),
    "lm_seq_gic": ("gaussian", 150, 200, 5, 2, False, 5, 4, 1, 10, 0.01, 100.0, 8, True, False, 63),
    "lm_seq_cv": ("gaussian", 150, 200, 5, 2, True, 3, 1, 2, 9, 0.01, 10.0, 6, True, False, 64),
    "lm_gs_cold": ("gaussian", 150, 200, 5, 1, False, 5, 2, 1, 12, 0.01, 100.0, 100, False, False, 65),
    "logit_gs_gic": ("binomial", 200, 150, 4, 1, False, 5, 3, 1, 10, 0.001, 1.0, 100, True, False, 66),
    "logit_seq_cv": ("binomial", 200, 150, 4, 2, True, 3, 1, 1, 8, 0.001, 1.0, 5, True, False, 67),
    "poisson_gs_cv": ("poisson", 200, 150, 4, 1, True, 3, 1, 1, 8, 0.001, 1.0, 100, True, False, 68),
    "poisson_seq_gic": ("poisson", 200, 150, 4, 2, False, 5, 3, 1, 5, 0.001, 1.0, 4, True, False, 69),
    "cox_gs_gic": ("cox", 160, 150, 4, 1, False, 5, 2, 1, 8, 0.001, 0.05, 100, True, False, 70),
    "cox_seq_cv": ("cox", 160, 150, 4, 2, True, 3, 1, 1, 7, 0.001, 0.05, 5, True, False, 71),
}


def main():
    only = set(sys.argv[1:])
    for name, (fam, n, p, k, pp, is_cv, K, ic_type, s_min, s_max, lmin, lmax, nl, warm, weighted, seed) in CASES.items():
        if only and name not in only:
            continue
        model_type, data_type = FAM[fam]
        d = gen_data(n, p, fam, k, seed=seed)
        rng = np.random.Generator(np.random.PCG64(2000 + seed))
        w = rng.uniform(0.5, 1.5, n) if weighted else np.ones(n)
        fold = ref.cv_fold_ids(n, K) if is_cv else np.zeros(n, dtype=np.int32)
        r = ref.bess_lambda(d.x, d.y, data_type, w, True, 5, model_type, 20, 2, warm, ic_type, is_cv, K, [1], s_min, s_max,
                            lambda_min=lmin, lambda_max=lmax, n_lambda=nl, powell_path=pp)
        t = ref.pgs_trace(d.x, d.y, data_type, w, True, 5, model_type, 20, warm, ic_type, is_cv, K, s_min, s_max, lmin, lmax,
                          nl, pp)
        # the traced run drives pgs_path directly; it must agree with the bessCpp run bit for bit
        assert np.array_equal(t["beta"], r["beta"]) and t["ic"] == r["ic"] and t["lambda_"] == r["lambda_"], name
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"), x=d.x, y=d.y, weight=w, fold_of_row=fold, beta=r["beta"], coef0=r["coef0"],
            train_loss=r["train_loss"], ic=r["ic"], lam=r["lambda_"], fits=t["fits"],
            meta=np.array([model_type, data_type, pp, int(is_cv), K, ic_type, s_min, s_max, nl, int(warm)], dtype=np.int64),
            lambda_range=np.array([lmin, lmax]))
        print(name, "support", np.nonzero(r["beta"])[0].tolist(), "ic", r["ic"], "lambda", r["lambda_"], "fits", t["n_fits"])


if __name__ == "__main__":
    main()
