import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bess_b200 import _lib
from bess_b200.engine import GpuEngine
from oracle import pdas_oracle as orc
from tests.helpers import load_golden, rel_err

g = load_golden("cox_gs_bic")
data = orc.make_data(g["x"], g["y"], g["weight"], 3, True, 4)
lib = _lib.load()
orig = orc.fit_cox
for N in (1, 2, 3, 5, 8, 12, 20, 30):
    lib.bess_b200_debug_set(1, N)
    def fit_n(XA, st, w, c, N=N):
        # oracle with the Newton loop truncated to N steps
        import math
        n, k = XA.shape
        beta0 = np.zeros(k); ll0 = 1e5; ws = w * st
        for _ in range(N):
            theta = np.exp(np.clip(XA @ beta0, -30, 30))
            cum = np.cumsum(theta[::-1])[::-1]
            xt = np.cumsum((XA * theta[:, None])[::-1], axis=0)[::-1] / cum[:, None]
            gg = (XA - xt).T @ ws
            h = np.empty((k, k))
            for a in range(k):
                for b in range(a, k):
                    s = np.cumsum((theta * XA[:, a] * XA[:, b])[::-1])[::-1]
                    h[a, b] = h[b, a] = -float((s / cum - xt[:, a] * xt[:, b]) @ ws)
            d = np.linalg.solve(h, gg)
            m = 1; beta1 = beta0 - 0.5 ** m * d; ll1 = orc.loglik_cox(XA, st, beta1, w)
            while ll0 > ll1 and m < 5:
                m += 1; beta1 = beta0 - 0.5 ** m * d; ll1 = orc.loglik_cox(XA, st, beta1, w)
            if abs(ll0 - ll1) / abs(0.1 + ll0) < 1e-5:
                break
            beta0 = beta1; ll0 = ll1
        return beta0, c
    orc._FIT[4] = fit_n
    eng = GpuEngine()
    eng.load(g["x"], g["y"], g["weight"], 4)
    eng.normalize(3, True)
    eng.setup_chains(0, None, 10, 20, True)
    binit = np.zeros(data.p)
    st = orc.PathState(data, 4, 2, False, 0, None, 20, True)
    for T in (2, 5, 8, 10):
        r = eng.run_batch(T, [0], True)
        o = orc.pdas_fit(data, 4, T, binit, 0.0, st.full_mask, None, 20)
        okA = np.array_equal(o.A, r["A"][0])
        print(f"N={N} T={T} A_ok={okA} l={r['l'][0]}/{o.l} beta_err={rel_err(r['bA'][0], o.beta[o.A]) if okA else -1:.2e} maxbeta={np.abs(o.beta).max():.3g}", flush=True)
        binit = o.beta
    eng.close()
