"""Roofline probe of the batched dual sweep: F chains in one pass over a resident 1000 x 500000 design (C5b shape).
Run alone for timings, or under ncu (-k regex:dual_sweep) for the full-section capture."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bess_b200 import cbess  # noqa: E402
from bess_b200.engine import GpuEngine  # noqa: E402


def main():
    n, p = 1000, int(os.environ.get("PROBE_P", "500000"))
    reps = int(os.environ.get("PROBE_REPS", "20"))
    g = torch.Generator(device="cuda").manual_seed(5)
    X = torch.randn(n, p, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(n, dtype=torch.float64, device="cuda", generator=g).cpu().numpy()
    w = np.ones(n)
    for K in [int(v) for v in os.environ.get("PROBE_K", "10,0").split(",")]:
        eng = GpuEngine()
        eng.load(None, y, w, 1, x_device_ptr=X.data_ptr(), n=n, p=p)
        eng.normalize(1, True)
        eng.setup_chains(K, cbess.cv_fold_ids(n, K, 123) if K else None, 20, 20, True)
        eng.run_batch(3, list(range(K + 1)), True)
        ms, by = eng.time_dual_sweep(reps)
        print(f"F={K + 1} chains: {ms:.4f} ms/launch  {by / ms / 1e6:.0f} GB/s algorithmic (8np)", flush=True)
        eng.close()


if __name__ == "__main__":
    main()
