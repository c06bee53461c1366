"""Host wall-clock of one C5 call by phase (no CUDA-event profiling), design resident in HBM: where the time outside the
kernels goes.  Prints the mean over the timed calls."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bess_b200 import _lib, cbess  # noqa: E402


def main():
    n, p, k = 1000, 500000, 10
    g = torch.Generator(device="cuda").manual_seed(5)
    X = torch.randn(n, p, dtype=torch.float64, device="cuda", generator=g)
    rng = np.random.default_rng(5)
    nz = np.sort(rng.choice(p, k, replace=False))
    beta = rng.uniform(1, 5, k)
    y = (X[:, torch.as_tensor(nz, device="cuda")] @ torch.as_tensor(beta, device="cuda")).cpu().numpy() + rng.normal(0, 3, n)
    w = np.ones(n)
    seq = np.arange(1, 21)
    for fg in [int(v) for v in os.environ.get("FIRST_GROUPS", "3").split(",")]:
        _lib.load().bess_b200_debug_set(3, fg)
        run(X, y, w, seq, n, p, fg)
    _lib.load().bess_b200_debug_set(3, 3)


def run(X, y, w, seq, n, p, fg):
    acc, wall, reps = None, 0.0, 30
    for r in range(reps + 3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = cbess.fit(None, y, 1, w, True, 1, 1, 20, 2, 1, True, 1, True, 10, seq, 1, 20, True, 5000,
                        x_device_ptr=X.data_ptr(), n=n, p=p, want_trace=False)
        dt = time.perf_counter() - t0
        if r >= 3:
            h = np.array(list(out["stats"]["host_ms"].values()))
            acc = h if acc is None else acc + h
            wall += dt
    names = list(out["stats"]["host_ms"].keys())
    print("first group", fg, "| C5 resident, mean of", reps, "calls: wall", round(wall / reps * 1e3, 3), "ms; host phases (ms):",
          {nm: round(v / reps, 3) for nm, v in zip(names, acc)}, "sum", round(acc.sum() / reps, 3),
          "| sweeps", out["stats"]["n_sweeps"], "batches", out["stats"]["n_batches"], "launches", out["stats"]["kernel_launches"])
    r = out["stats"]["resident"]
    if r[0] > 0:
        mhz = 1965.0
        own = dict(zip(("begin", "wait", "select", "load_cols", "gram", "solve", "resid_cycle", "publish"), [round(v / mhz, 1) for v in r[8:16]]))
        swp = dict(zip(("wait", "stream_x", "reduce_sacrifice"), [round(v / mhz, 1) for v in r[16:19]]))
        print("  resident kernel (last call): launches", r[0], "iterations", r[1], "fallback selects", r[2], "steps", r[3], "merged level starts", r[4], "| assemble, cand load, slots, scatter+cycle, bookkeeping us", [round(v / mhz, 1) for v in r[19:24]],
              "| owner 0 us by phase", own, "| sweeper 0 us", swp)
        print("  DEBUG chol: factor us (cumulative over calls)", round(r[5] / mhz, 1), "backsub us", round(r[6] / mhz, 1), "avg k", round(r[7] / mhz, 2))
        ow = np.array(r[24:152]).reshape(32, 4) / mhz
        print("  per owner: busy us", ow[:11, 0].round(0).tolist(), "| longest phase us", ow[:11, 1].round(1).tolist(),
              "| fallback us", ow[:11, 2].round(0).tolist(), "| fits solved", (ow[:11, 3] * mhz).round(0).tolist())


if __name__ == "__main__":
    main()
