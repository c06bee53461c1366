#!/bin/bash
# ncu launch list (gpu__time_duration) of two C5 calls (resident design): which kernels make up a step
mkdir -p gpurun_out
cat > /tmp/two_calls.py <<'PY'
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
print(out["s"], out["stats"]["kernel_launches"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5.csv python /tmp/two_calls.py > gpurun_out/ncu_list.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_c5.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
names=[(r[ki][:70], float(r[vi].replace(',','')), r[ui]) for r in rows[1:]]
# last call = last third
for nm,v,u in names[-(len(names)//3+2):]:
    print(f"{v/1000 if u.startswith('n') else v:10.1f} us  {nm}")
PY
