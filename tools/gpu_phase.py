"""chain_fit phase breakdown (clock64 timers inside the kernel, thread 0 of each cluster's rank 0) for one BASELINE config."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bess_b200 import _lib, cbess  # noqa: E402
from bess_b200.gen_data import gen_data  # noqa: E402
from tests.helpers import FULL_CONFIGS  # noqa: E402

FAM = {"gaussian": (1, 1), "binomial": (2, 2), "poisson": (3, 2), "cox": (4, 3)}
NAMES = ["gather", "eval", "wz", "syrk", "reduce", "chol/back", "bcast", "grad", "cycle", "other", "c:load", "c:diag", "c:rows",
         "c:wb+syncA", "c:upd+syncB", "dot"]


def main():
    lib = _lib.load()
    for cfg in sys.argv[1:] or ["c2"]:
        fam, n, p, k, path_type, is_cv, K, ic_type, s_min, s_max, scr, seed = FULL_CONFIGS[cfg]
        model_type, data_type = FAM[fam]
        d = gen_data(n, p, fam, k, seed=seed)
        w = np.ones(n)
        seq = np.arange(s_min, s_max + 1) if path_type == 1 else np.arange(1, 2)
        for rep in range(2):
            lib.bess_b200_debug_set(2, 1)
            t0 = time.time()
            out = cbess.fit(d.x, d.y, data_type, w, True, 1, model_type, 20, 2, path_type, True, ic_type, is_cv, K, seq,
                            s_min, s_max, scr > 0, max(scr, 1), cv_seed=123, want_trace=False, profile=True)
            dt = time.time() - t0
            buf = (C.c_ulonglong * 32)()
            lib.bess_b200_debug_get(buf)
            lib.bess_b200_debug_set(2, 0)
        ticks, hits = np.array(buf[:16], dtype=np.float64), np.array(buf[16:], dtype=np.float64)
        tot = ticks.sum()
        print(f"{cfg}: call {dt * 1e3:.1f} ms, chain kernels {out['stats']['prof_ms']['chain']:.1f} ms, fits {out['stats']['n_fits']}, "
              f"pdas iters {out['stats']['n_pdas_iters']}; phase timers summed over chains (1.965 GHz):")
        for i, nm in enumerate(NAMES):
            if hits[i] > 0:
                print(f"   {nm:8s} {ticks[i] / 1.965e6:10.2f} ms  {100 * ticks[i] / tot:5.1f}%  hits {int(hits[i]):7d}  "
                      f"{ticks[i] / hits[i] / 1965:8.2f} us/hit")


if __name__ == "__main__":
    main()
