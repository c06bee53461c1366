"""GPU bring-up script (not a test): compares every layer of the CUDA path with the numpy oracle and prints diffs."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bess_b200 import cbess  # noqa: E402
from bess_b200.engine import GpuEngine, topk  # noqa: E402
from bess_b200.gen_data import gen_data  # noqa: E402
from oracle import pdas_oracle as orc  # noqa: E402
from tests.helpers import golden_names, load_golden, rel_err  # noqa: E402

FAM = {"gaussian": (1, 1), "binomial": (2, 2), "poisson": (3, 2), "cox": (4, 3)}


def section(t):
    print("\n==== " + t, flush=True)


def check_topk():
    section("topk")
    rng = np.random.default_rng(0)
    for n in (100, 5000, 16384, 20000, 100000, 500000):
        for k in (1, 5, 20, 263, 5000):
            if k > n:
                continue
            v = rng.random(n) ** 4
            if n == 5000:
                v[rng.integers(0, n, 50)] = np.finfo(np.float64).max
            exp = orc.max_k(v, k)
            got, tie = topk(v, k)
            ok = np.array_equal(exp, got)
            print(f"n={n} k={k} ok={ok} tie={tie}", flush=True)
            if not ok:
                print("   exp", exp[:10], "got", got[:10])
    v = np.floor(rng.random(3000) * 10)
    got, tie = topk(v, 500)
    print("ties: ok=", np.array_equal(orc.max_k(v, 500), got), "tie flag", tie)


def check_batches(fam, n, p, k, K, Ts, seed=3, weighted=False):
    section(f"batches {fam} n={n} p={p} K={K}")
    model_type, data_type = FAM[fam]
    d = gen_data(n, p, fam, k, seed=seed)
    rng = np.random.default_rng(seed)
    w = rng.uniform(0.5, 1.5, n) if weighted else np.ones(n)
    fold = cbess.cv_fold_ids(n, K, 123) if K else None
    eng = GpuEngine()
    eng.load(d.x, d.y, w, model_type)
    xm, xn, ym = eng.normalize(data_type, True)
    data = orc.make_data(d.x, d.y, w, data_type, True, model_type)
    print("xmean err", rel_err(xm, data.x_mean) if data_type != 3 else 0.0, "xnorm err", rel_err(xn, data.x_norm),
          "ymean", ym, data.y_mean)
    eng.setup_chains(K, fold, max(Ts), 20, True)
    st = orc.PathState(data, model_type, 3, K > 0, K, fold, 20, True)
    chains = list(range(K + 1))
    binit = [np.zeros(p) for _ in chains]
    c0_level = 0.0
    c0_full = 0.0
    for T in Ts:
        t0 = time.time()
        r = eng.run_batch(T, chains, True)
        dt = time.time() - t0
        c0_level = c0_full
        masks = [st.full_mask] + (st.train_masks if K else [])
        xtxs = [st.xtx_full] + (st.xtx_folds if K else [])
        for ci in chains:
            o = orc.pdas_fit(data, model_type, T, binit[ci], c0_level, masks[ci], xtxs[ci], 20)
            okA = np.array_equal(o.A, r["A"][ci])
            eb = rel_err(r["bA"][ci], o.beta[o.A]) if okA else float("nan")
            print(f"T={T} chain={ci} A_ok={okA} l={r['l'][ci]}/{o.l} beta_err={eb:.2e} coef0={r['coef0'][ci]:.10g}/{o.coef0:.10g}"
                  + (f" gap={o.min_gap:.1e}" if not okA else ""), flush=True)
            if not okA:
                print("    gpu A", r["A"][ci][:12], "\n    orc A", o.A[:12], "hist", [a[:6].tolist() for a in o.A_hist][:3])
            binit[ci] = o.beta
            if ci == 0:
                c0_full = o.coef0
        # losses
        jobs = [(0, 0, 0)] + [(1 + kk, 1, kk) for kk in range(K)]
        lv = eng.losses(jobs)
        exp = [orc.train_loss(data, model_type, binit[0], c0_full)]
        for kk in range(K):
            exp.append(orc.fold_loss(data, model_type, binit[1 + kk], r["coef0"][1 + kk], st.test_masks[kk]))
        print(f"   losses err {rel_err(lv, np.array(exp)):.2e}  batch {dt*1e3:.2f} ms", flush=True)
    print("stats", eng.stats())
    eng.close()


def check_golden():
    section("golden end-to-end")
    for name in golden_names():
        g = load_golden(name)
        seq = np.arange(1, g["smax"] + 1)
        try:
            t0 = time.time()
            out = cbess.fit(g["x"], g["y"], g["data_type"], g["weight"], True, 1, g["model_type"], 20, 2, g["path_type"],
                            True, g["ic_type"], g["is_cv"], g["K"], seq, 1, g["smax"], g["scr"] > 0, max(g["scr"], 1),
                            fold_of_row=g["fold_of_row"] if g["is_cv"] else None)
            dt = time.time() - t0
            sa = np.nonzero(out["beta"])[0].tolist()
            sb = np.nonzero(g["beta"])[0].tolist()
            print(f"{name}: support_ok={sa == sb} beta_err={rel_err(out['beta'], g['beta']):.2e} "
                  f"coef0 {out['coef0']:.10g}/{g['coef0']:.10g} loss_err={abs(out['train_loss']-g['train_loss'])/abs(g['train_loss']):.2e} "
                  f"ic_err={abs(out['ic']-g['ic'])/abs(g['ic']):.2e} s={out['s']} {dt*1e3:.1f} ms ties={out['stats']['n_boundary_ties']}",
                  flush=True)
            if "screening_A" in g:
                print("   screening_A ok:", np.array_equal(out["screening_A"], g["screening_A"]))
            if sa != sb:
                print("   gpu", sa, "ref", sb)
        except Exception as e:
            print(name, "FAILED:", e)
            traceback.print_exc()


def main():
    which = sys.argv[1:] or ["topk", "lm", "logit", "poisson", "cox", "golden"]
    for w in which:
        try:
            if w == "topk":
                check_topk()
            elif w == "lm":
                check_batches("gaussian", 300, 1000, 8, 3, [1, 2, 5, 9], weighted=True)
            elif w == "logit":
                check_batches("binomial", 400, 600, 5, 3, [1, 3, 6])
            elif w == "poisson":
                check_batches("poisson", 400, 600, 5, 2, [1, 3, 6])
            elif w == "cox":
                check_batches("cox", 300, 500, 5, 2, [1, 3, 6])
            elif w == "golden":
                check_golden()
        except Exception as e:
            print("SECTION", w, "FAILED:", e)
            traceback.print_exc()


if __name__ == "__main__":
    main()
