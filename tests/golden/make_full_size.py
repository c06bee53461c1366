"""Full-size goldens for the five BASELINE configs, produced by the REAL reference (oracle/_ref/libbess_ref.so).

    python tests/golden/make_full_size.py c1 c1cv c2 c3 c4 c5      (build container only; C2-C4 take 10-20 CPU-minutes each)

The designs are NOT stored (up to 4 GB): they are regenerated bit-identically from the seed by bess_b200.gen_data
(numpy PCG64) wherever the test runs.  Each tests/golden/full/<cfg>.npz keeps the reference's outputs in sparse form
(support, coefficients on it, coef0, train_loss, ic), the CV fold assignment it drew (seed 123), a checksum of the
inputs, and the single-thread wall time of the reference call on this container's CPU."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bess_b200.gen_data import gen_data  # noqa: E402
from oracle import ref  # noqa: E402
from tests.helpers import FULL_CONFIGS, full_checksum as checksum  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FAM = {"gaussian": (1, 1), "binomial": (2, 2), "poisson": (3, 2), "cox": (4, 3)}

def main():
    for name in sys.argv[1:]:
        fam, n, p, k, path_type, is_cv, K, ic_type, s_min, s_max, scr, seed = FULL_CONFIGS[name]
        model_type, data_type = FAM[fam]
        d = gen_data(n, p, fam, k, seed=seed)
        w = np.ones(n)
        seq = np.arange(s_min, s_max + 1, dtype=np.int32) if path_type == 1 else np.arange(1, 2, dtype=np.int32)
        fold = ref.cv_fold_ids(n, K, 123) if is_cv else np.zeros(n, dtype=np.int32)
        t0 = time.perf_counter()
        r = ref.pywrap_bess(d.x, d.y, data_type, w, True, 1, model_type, 20, 2, path_type, True, ic_type, is_cv, K, seq,
                            s_min, s_max, scr > 0, scr if scr > 0 else 1, cv_seed=123)
        dt = time.perf_counter() - t0
        sup = np.nonzero(r["beta"])[0].astype(np.int64)
        out = dict(support=sup, beta_support=r["beta"][sup], coef0=r["coef0"], train_loss=r["train_loss"], ic=r["ic"],
                   fold_of_row=fold, checksum=checksum(d), ref_seconds=dt, true_support=np.nonzero(d.beta)[0])
        if scr > 0:
            out["screening_A"] = ref.screening(d.x, d.y, w, model_type, scr)
        np.savez_compressed(os.path.join(OUT, "full", f"{name}.npz"), **out)
        print(name, "reference seconds", round(dt, 2), "s =", sup.size, "support", sup[:12].tolist(), "ic", r["ic"], flush=True)


if __name__ == "__main__":
    main()
