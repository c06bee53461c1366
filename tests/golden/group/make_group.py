"""Generates tests/golden/group/*.npz from the REAL reference (oracle/_ref/libbess_ref.so): group selection (gsize > 1),
i.e. bessCpp with a g_index shorter than p and algorithm_type 2 (GPDAS) / 3 (GL0L2) -- /root/reference/src/Algorithm.h
:1097-1129, 1206-1263, 1324-1367, 1497-1568 and utilities.cpp:113-177.  Sparsity levels and always_select count groups.
Run in the build container only:  python tests/golden/group/make_group.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
from bess_b200.gen_data import gen_data  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FAM = {"gaussian": (1, 1), "binomial": (2, 2), "poisson": (3, 2), "cox": (4, 3)}


def layout(p, sizes):
    """first column of every group, group sizes cycling through `sizes`; the last group takes what is left."""
    gi, k = [0], 0
    while gi[-1] + sizes[k % len(sizes)] < p:
        gi.append(gi[-1] + sizes[k % len(sizes)])
        k += 1
    return np.array(gi, dtype=np.int32)


# name: family, n, p, k, sizes, algorithm_type, path_type, is_cv, K, ic_type, s_min, s_max, lambdas / (lmin, lmax, nl, powell),
#       weighted, always (group numbers), seed
CASES = {
    "lm_seq_gic": ("gaussian", 150, 200, 6, [3, 1, 4, 2], 2, 1, False, 5, 3, 1, 8, [0.0], False, (), 101),
    "lm_seq_cv_w": ("gaussian", 160, 180, 6, [2, 5, 1], 2, 1, True, 4, 1, 1, 7, [0.0], True, (), 102),
    "lm_gs_cv": ("gaussian", 150, 200, 6, [3, 1, 4, 2], 2, 2, True, 3, 1, 1, 12, [0.0], False, (), 103),
    "lm_wide_always": ("gaussian", 200, 240, 6, [8, 3, 6, 1, 7], 2, 1, False, 5, 4, 2, 6, [0.0], False, (2, 11), 104),
    "lm_l0l2_seq": ("gaussian", 150, 200, 6, [3, 1, 4, 2], 3, 1, False, 5, 3, 1, 6, [0.0, 0.05, 0.5], False, (), 105),
    "lm_l0l2_pgs": ("gaussian", 150, 200, 6, [3, 1, 4, 2], 3, 2, False, 5, 3, 1, 8, (0.01, 10.0, 100, 1), False, (), 106),
    "logit_seq_cv": ("binomial", 220, 160, 4, [2, 3, 1], 2, 1, True, 3, 1, 1, 6, [0.0], False, (), 107),
    "logit_gs_bic": ("binomial", 220, 160, 4, [4, 1, 2], 2, 2, False, 5, 2, 1, 9, [0.0], False, (), 108),
    "logit_l0l2_seq": ("binomial", 220, 160, 4, [2, 3, 1], 3, 1, False, 5, 3, 1, 5, [0.01, 0.1], False, (), 109),
    "poisson_seq_gic": ("poisson", 220, 160, 4, [3, 1, 2], 2, 1, False, 5, 3, 1, 6, [0.0], False, (), 110),
    "poisson_seq_cv": ("poisson", 220, 160, 4, [2, 2, 4], 2, 1, True, 3, 1, 1, 5, [0.0], False, (), 111),
    "cox_seq_gic": ("cox", 180, 150, 4, [3, 1, 4, 2], 2, 1, False, 5, 3, 1, 6, [0.0], False, (), 112),
    "cox_seq_cv": ("cox", 180, 150, 4, [2, 3], 2, 1, True, 3, 1, 1, 5, [0.0], False, (), 113),
    "cox_l0l2_seq": ("cox", 180, 150, 4, [3, 1, 4, 2], 3, 1, False, 5, 2, 1, 5, [0.0, 0.01], False, (3,), 114),
    # every group a single variable (GroupPdas* with group = 0..p-1): algorithm_type 2 / 3 take the group code paths of the
    # reference -- for cox the dense-Hessian branch, Algorithm.h:1497-1568 -- which must agree with the plain PDAS build
    "single_cox_seq_cv": ("cox", 160, 120, 4, [1], 2, 1, True, 3, 1, 1, 6, [0.0], False, (), 115),
    "single_lm_gs_cv": ("gaussian", 150, 200, 5, [1], 2, 2, True, 3, 1, 1, 12, [0.0], False, (), 116),
    "single_logit_l0l2_seq": ("binomial", 200, 150, 4, [1], 3, 1, False, 5, 3, 1, 6, [0.01, 0.1], False, (), 117),
    # groups wider than 8 variables (the shared-memory sacrifice kernel; utilities.cpp:142-177 takes any size)
    "wide_lm_seq_gic": ("gaussian", 200, 240, 6, [12, 3, 1, 17, 5], 2, 1, False, 5, 3, 1, 5, [0.0], False, (), 121),
    "wide_lm_l0l2_cv": ("gaussian", 220, 200, 5, [9, 2, 20, 4], 3, 1, True, 3, 1, 1, 4, [0.0, 0.1], True, (), 122),
    "wide_logit_seq_gic": ("binomial", 400, 160, 4, [10, 2, 13, 1], 2, 1, False, 5, 3, 1, 4, [0.0], False, (), 123),
    "wide_poisson_seq_cv": ("poisson", 400, 150, 4, [11, 3, 2], 2, 1, True, 3, 1, 1, 4, [0.0], False, (), 124),
    "wide_cox_seq_gic": ("cox", 300, 150, 4, [12, 1, 9, 3], 2, 1, False, 5, 3, 1, 4, [0.0], False, (), 125),
    "wide_cox_l0l2_cv": ("cox", 300, 140, 4, [10, 4, 16], 3, 1, True, 3, 1, 1, 3, [0.0, 0.01], False, (1,), 126),
}


def main():
    only = set(sys.argv[1:])
    for name, (fam, n, p, k, sizes, alg, path_type, is_cv, K, ic_type, s_min, s_max, lam, weighted, always, seed) in CASES.items():
        if only and name not in only:
            continue
        model_type, data_type = FAM[fam]
        d = gen_data(n, p, fam, k, seed=seed)
        rng = np.random.Generator(np.random.PCG64(3000 + seed))
        w = rng.uniform(0.5, 1.5, n) if weighted else np.ones(n)
        gi = layout(p, sizes)
        fold = ref.cv_fold_ids(n, K) if is_cv else np.zeros(n, dtype=np.int32)
        pgs = isinstance(lam, tuple)
        seq = np.arange(s_min, s_max + 1, dtype=np.int32)
        kw = dict(lambda_min=lam[0], lambda_max=lam[1], n_lambda=lam[2], powell_path=lam[3]) if pgs else dict(lambda_seq=lam)
        r = ref.bess_lambda(d.x, d.y, data_type, w, True, alg, model_type, 20, path_type, True, ic_type, is_cv, K, seq, s_min,
                            s_max, always_select=always, g_index=gi, **kw)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"), x=d.x, y=d.y, weight=w, fold_of_row=fold, g_index=gi, beta=r["beta"],
            coef0=r["coef0"], train_loss=r["train_loss"], ic=r["ic"], lam=r["lambda_"],
            always=np.array(always, dtype=np.int32),
            lambdas=np.array(lam[:2] if pgs else lam, dtype=np.float64),
            meta=np.array([model_type, data_type, alg, path_type, int(is_cv), K, ic_type, s_min, s_max, int(pgs),
                           lam[2] if pgs else 0, lam[3] if pgs else 1], dtype=np.int64))
        print(name, "groups", len(gi), "support", np.nonzero(r["beta"])[0].tolist(), "ic", r["ic"], "lambda", r["lambda_"])


if __name__ == "__main__":
    main()
