#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/two_calls.py <<PY
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from bess_b200 import cbess
n, p, k = 1000, 500000, 10
g = torch.Generator(device="cuda").manual_seed(5)
X = torch.randn(n, p, dtype=torch.float64, device="cuda", generator=g)
rng = np.random.default_rng(5)
nz = np.sort(rng.choice(p, k, replace=False))
beta = rng.uniform(1, 5, k)
y = (X[:, torch.as_tensor(nz, device="cuda")] @ torch.as_tensor(beta, device="cuda")).cpu().numpy() + rng.normal(0, 3, n)
w = np.ones(n); seq = np.arange(1, 21)
for r in range(3):
    out = cbess.fit(None, y, 1, w, True, 1, 1, 20, 2, 1, True, 1, True, 10, seq, 1, 20, True, 5000, x_device_ptr=X.data_ptr(), n=n, p=p, want_trace=False)
print(out["s"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_path_kernel -s 1 -c 1 -o gpurun_out/prof_lm_path -f python /tmp/two_calls.py > gpurun_out/ncu_lp.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_lp.log; ls -la gpurun_out/prof_lm_path.ncu-rep
