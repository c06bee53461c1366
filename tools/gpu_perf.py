"""GPU exploration: times the BASELINE configs through the C-ABI and prints the per-category device-time profile."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bess_b200 import cbess  # noqa: E402
from bess_b200.engine import GpuEngine  # noqa: E402
from bess_b200.gen_data import gen_data  # noqa: E402

FAM = {"gaussian": (1, 1), "binomial": (2, 2), "poisson": (3, 2), "cox": (4, 3)}


def show(tag, out, dt):
    st = out["stats"]
    prof = {k: round(v, 3) for k, v in st["prof_ms"].items()}
    print(f"{tag}: {dt*1e3:.1f} ms  s={out['s']} support={np.nonzero(out['beta'])[0][:12].tolist()} fits={st['n_fits']} "
          f"iters={st['n_pdas_iters']} sweeps={st['n_sweeps']} launches={st['kernel_launches']} ties={st['n_boundary_ties']} "
          f"S={st['sweep_splits']}\n    prof_ms={prof} n={st['prof_launches']}\n    host_ms={ {k: round(v, 3) for k, v in st['host_ms'].items()} }",
          flush=True)
    if st["prof_ms"]["dual_sweep"] > 0:
        print(f"    PDAS sweep: {st['sweep_bytes']/st['prof_ms']['dual_sweep']/1e6:.0f} GB/s algorithmic; "
              f"screening sweep: {st['big_sweep_bytes']/max(st['prof_ms']['screen_sweep'],1e-9)/1e6:.0f} GB/s", flush=True)


def run(tag, fam, n, p, k, path_type, is_cv, K, ic_type, seq, s_min, s_max, scr, seed, reps=2, x=None, d=None):
    model_type, data_type = FAM[fam]
    if d is None:
        d = gen_data(n, p, fam, k, seed=seed)
    w = np.ones(n)
    for r in range(reps):
        t0 = time.time()
        out = cbess.fit(d.x, d.y, data_type, w, True, 1, model_type, 20, 2, path_type, True, ic_type, is_cv, K, seq,
                        s_min, s_max, scr > 0, max(scr, 1), profile=True, want_trace=False)
        dt = time.time() - t0
        show(f"{tag} rep{r}", out, dt)
    print("    true support", np.nonzero(d.beta)[0].tolist())
    return d


def c5(resident=True):
    n, p, k = 1000, 500000, 10
    g = torch.Generator(device="cuda").manual_seed(5)
    X = torch.randn(n, p, dtype=torch.float64, device="cuda", generator=g)
    rng = np.random.default_rng(5)
    nz = np.sort(rng.choice(p, k, replace=False))
    m = 5 * np.sqrt(2 * np.log(p) / n)
    beta = rng.uniform(m, 100 * m, k)
    sigma = np.sqrt((beta @ beta) / 10)
    y = (X[:, torch.as_tensor(nz, device="cuda")] @ torch.as_tensor(beta, device="cuda")).cpu().numpy() + rng.normal(0, sigma, n)
    w = np.ones(n)
    seq = np.arange(1, 21)
    for r in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        out = cbess.fit(None, y, 1, w, True, 1, 1, 20, 2, 1, True, 1, True, 10, seq, 1, 20, True, 5000,
                        x_device_ptr=X.data_ptr(), n=n, p=p, profile=True, want_trace=False)
        dt = time.time() - t0
        show(f"C5 resident rep{r}", out, dt)
    print("    true", nz.tolist())
    # e2e from pinned host
    Xh = torch.empty((n, p), dtype=torch.float64, pin_memory=True)
    Xh.copy_(X)
    xh = Xh.numpy()
    for r in range(2):
        t0 = time.time()
        out = cbess.fit(xh, y, 1, w, True, 1, 1, 20, 2, 1, True, 1, True, 10, seq, 1, 20, True, 5000, profile=True,
                        want_trace=False)
        dt = time.time() - t0
        show(f"C5 e2e(pinned host) rep{r}", out, dt)
    # C5b: no screening, sweeps are p=500k sized.  roofline probe
    eng = GpuEngine()
    eng.load(None, y, w, 1, x_device_ptr=X.data_ptr(), n=n, p=p)
    t0 = time.time(); eng.normalize(1, True); print("normalize 500k:", time.time() - t0)
    fold = cbess.cv_fold_ids(n, 10, 123)
    t0 = time.time(); eng.setup_chains(10, fold, 20, 20, True); print("setup_chains:", time.time() - t0)
    r = eng.run_batch(1, list(range(11)), True)
    ms, by = eng.time_dual_sweep(20)
    print(f"C5b dual sweep F=12 slots: {ms:.3f} ms/launch, {by/ms/1e6:.0f} GB/s algorithmic (8np)")
    t0 = time.time()
    for T in range(2, 6):
        r = eng.run_batch(T, list(range(11)), True)
    print("C5b 4 levels x 11 chains:", time.time() - t0, "s", eng.stats())
    eng.close()
    eng = GpuEngine()
    eng.load(None, y, w, 1, x_device_ptr=X.data_ptr(), n=n, p=p)
    eng.normalize(1, True)
    eng.setup_chains(0, None, 20, 20, True)
    eng.run_batch(1, [0], True)
    ms, by = eng.time_dual_sweep(20)
    print(f"C5b dual sweep F=1: {ms:.3f} ms/launch, {by/ms/1e6:.0f} GB/s algorithmic (8np)")
    eng.close()


def main():
    which = sys.argv[1:] or ["c1", "c5", "c4", "c2", "c3"]
    for wname in which:
        print("\n=====", wname, flush=True)
        if wname == "c1":
            run("C1", "gaussian", 500, 1000, 10, 1, False, 5, 3, np.arange(1, 21), 1, 20, 0, 1, reps=3)
            run("C1cv", "gaussian", 500, 1000, 10, 1, True, 10, 1, np.arange(1, 21), 1, 20, 0, 1, reps=2)
        elif wname == "c5":
            c5()
        elif wname == "c2":
            run("C2", "binomial", 2000, 20000, 20, 2, True, 10, 1, np.arange(1, 2), 1, 263, 0, 2, reps=2)
        elif wname == "c3":
            run("C3", "poisson", 5000, 50000, 30, 1, False, 5, 3, np.arange(1, 41), 1, 40, 0, 3, reps=2)
        elif wname == "c4":
            run("C4", "cox", 2000, 10000, 15, 1, True, 5, 1, np.arange(1, 31), 1, 30, 0, 4, reps=2)


if __name__ == "__main__":
    main()
