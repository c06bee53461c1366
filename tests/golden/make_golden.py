"""Generates tests/golden/*.npz from the REAL reference (oracle/_ref/libbess_ref.so, built by oracle/Makefile from
/root/reference/src).  Run in the build container only:  python tests/golden/make_golden.py
The reference ships no golden vectors, so these are its outputs on seeded synthetic inputs.
Each file holds the inputs, the CV fold assignment the reference drew (seed pinned by oracle/ref_shim.h), the final
outputs of pywrap_bess, and -- for sequential paths -- the per-level trace from ref_seq_trace."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bess_b200.gen_data import gen_data  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FAM = {"gaussian": (1, 1), "binomial": (2, 2), "poisson": (3, 2), "cox": (4, 3)}  # family -> (model_type, data_type)

CASES = [
    # name, family, n, p, k, path_type, is_cv, K, ic_type, seq_max/s_max, screening_size, weighted, seed
    ("lm_seq_gic", "gaussian", 150, 300, 5, 1, False, 5, 3, 10, 0, False, 11),
    ("lm_seq_cv", "gaussian", 150, 300, 5, 1, True, 5, 1, 10, 0, False, 12),
    ("lm_gs_cv_w", "gaussian", 160, 240, 6, 2, True, 4, 1, 14, 0, True, 13),
    ("lm_seq_ebic_screen", "gaussian", 120, 800, 4, 1, False, 5, 4, 8, 100, False, 14),
    ("logit_seq_gic", "binomial", 200, 250, 4, 1, False, 5, 3, 8, 0, False, 21),
    ("logit_gs_cv", "binomial", 200, 250, 4, 2, True, 4, 1, 12, 0, False, 22),
    ("logit_seq_cv_w", "binomial", 180, 200, 4, 1, True, 3, 1, 7, 0, True, 23),
    ("poisson_seq_gic", "poisson", 200, 250, 4, 1, False, 5, 3, 8, 0, False, 31),
    ("poisson_gs_cv", "poisson", 200, 250, 4, 2, True, 4, 1, 12, 0, False, 32),
    ("cox_seq_cv", "cox", 160, 200, 4, 1, True, 3, 1, 7, 0, False, 41),
    ("cox_gs_bic", "cox", 160, 200, 4, 2, False, 5, 2, 10, 0, False, 42),
    ("logit_seq_screen", "binomial", 150, 400, 3, 1, False, 5, 3, 6, 60, False, 24),
    ("poisson_seq_screen", "poisson", 150, 300, 3, 1, False, 5, 3, 6, 50, False, 33),
    ("cox_seq_screen", "cox", 120, 300, 3, 1, False, 5, 3, 6, 50, False, 43),
]
# L0L2 ("bsrr", Algorithm::lambda_level) on the sequential path: lambda grid walked zig-zag per level (path.cpp:50)
LAMBDA_CASES = {
    "lm_seq_l0l2": ("gaussian", 150, 300, 5, 1, False, 5, 3, 8, 0, False, 51, [0.0, 0.05, 0.5]),
    "lm_seq_l0l2_cv": ("gaussian", 150, 300, 5, 1, True, 3, 1, 6, 0, True, 52, [0.3, 0.01]),
    "logit_seq_l0l2_cv": ("binomial", 200, 250, 4, 1, True, 3, 1, 5, 0, False, 53, [0.02, 0.2]),
    "poisson_seq_l0l2": ("poisson", 200, 250, 4, 1, False, 5, 3, 6, 0, False, 54, [0.01, 0.1]),
    "cox_seq_l0l2": ("cox", 160, 200, 4, 1, False, 5, 3, 6, 0, False, 55, [0.0, 0.01, 0.05]),
}


def main():
    only = set(sys.argv[1:])
    cases = [c + ([0.0],) for c in CASES] + [(nm,) + v for nm, v in LAMBDA_CASES.items()]
    for (name, fam, n, p, k, path_type, is_cv, K, ic_type, smax, scr, weighted, seed, lams) in cases:
        if only and name not in only:
            continue
        model_type, data_type = FAM[fam]
        d = gen_data(n, p, fam, k, seed=seed)
        rng = np.random.Generator(np.random.PCG64(1000 + seed))
        w = rng.uniform(0.5, 1.5, n) if weighted else np.ones(n)
        seq = np.arange(1, smax + 1, dtype=np.int32)
        fold = ref.cv_fold_ids(n, K) if is_cv else np.zeros(n, dtype=np.int32)
        r = ref.pywrap_bess(d.x, d.y, data_type, w, True, 1, model_type, 20, 2, path_type, True, ic_type, is_cv, K,
                            seq, 1, smax, scr > 0, scr if scr > 0 else 1, lambda_seq=lams)
        out = dict(x=d.x, y=d.y, weight=w, fold_of_row=fold, beta=r["beta"], coef0=r["coef0"],
                   train_loss=r["train_loss"], ic=r["ic"],
                   meta=np.array([model_type, data_type, path_type, int(is_cv), K, ic_type, smax, scr], dtype=np.int64))
        if len(lams) > 1 or lams[0] != 0.0:
            out["lambda_seq"] = np.asarray(lams, dtype=np.float64)
        if scr > 0:
            out["screening_A"] = ref.screening(d.x, d.y, w, model_type, scr)
        elif path_type == 1 and "lambda_seq" not in out:
            t = ref.seq_trace(d.x, d.y, w, data_type, True, model_type, 20, True, ic_type, is_cv, K, seq)
            out.update({k2: v for k2, v in t.items()})
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "support", np.nonzero(r["beta"])[0].tolist(), "ic", r["ic"])


if __name__ == "__main__":
    main()
