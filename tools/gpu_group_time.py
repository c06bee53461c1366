"""GPU time of group-selection calls (n = 1000, p = 8000, groups of 4; n = 1000, p = 4000 logistic) with the chain-batched
sacrifice kernel against the per-chain warp kernel (BESS_B200_GROUP_PER_CHAIN=1 in a separate process)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bess_b200 import cbess  # noqa: E402
from bess_b200.gen_data import gen_data  # noqa: E402

FAM = {"gaussian": (1, 1), "binomial": (2, 2), "poisson": (3, 2)}
for fam, n, p, k, gs, is_cv, K, smax in (("gaussian", 1000, 8000, 12, 4, True, 5, 10), ("binomial", 1000, 4000, 8, 4, True, 5, 8),
                                         ("gaussian", 1000, 8000, 12, 8, True, 10, 6), ("poisson", 1000, 4000, 6, 2, True, 5, 6)):
    mt, dt = FAM[fam]
    d = gen_data(n, p, fam, k, seed=7)
    gi = np.arange(0, p, gs, dtype=np.int32)
    seq = np.arange(1, smax + 1, dtype=np.int32)
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        out = cbess.fit(d.x, d.y, dt, np.ones(n), True, 2, mt, 20, 2, 1, True, 1, is_cv, K, seq, 1, smax, False, 1, cv_seed=123,
                        g_index=gi, want_trace=False, profile=True)
        best = min(best, time.perf_counter() - t0) if best else time.perf_counter() - t0
    st = out["stats"]
    print(f"{fam} n={n} p={p} groups of {gs} K={K}: call {best * 1e3:.1f} ms, sweep category {st['prof_ms']['dual_sweep']:.2f} ms over "
          f"{st['prof_launches']['dual_sweep']} launches, chain {st['prof_ms']['chain']:.1f} ms, ic {out['ic']:.6f} s {out['s']}", flush=True)
