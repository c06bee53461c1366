"""Where the time of one C5 call goes outside the library: Python wrapper vs the C call itself."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bess_b200 import cbess, _lib
n, p, k = 1000, 500000, 10
from bess_b200.gen_data import gen_design_device
X = gen_design_device(n, p, 0.0, 5, 0)
rng = np.random.default_rng(5)
nz = np.sort(rng.choice(p, k, replace=False))
beta = rng.uniform(1, 5, k)
y = (X[:, torch.as_tensor(nz, device="cuda")] @ torch.as_tensor(beta, device="cuda")).cpu().numpy() + rng.normal(0, 3, n)
w = np.ones(n); seq = np.arange(1, 21)
lib = _lib.load()
orig = lib.bess_b200_fit
acc = {"c": 0.0, "n": 0}
class Wrap:
    def __call__(self, *a):
        t0 = time.perf_counter(); r = orig(*a); acc["c"] += time.perf_counter() - t0; acc["n"] += 1; return r
cbess_lib_fit = Wrap()
import bess_b200.cbess as cb
class LibProxy:
    def __getattr__(self, nm):
        return cbess_lib_fit if nm == "bess_b200_fit" else getattr(lib, nm)
_lib_load = _lib.load
_lib.load = lambda: LibProxy()
for r in range(5):
    cbess.fit(None, y, 1, w, True, 1, 1, 20, 2, 1, True, 1, True, 10, seq, 1, 20, True, 5000, x_device_ptr=X.data_ptr(), n=n, p=p, want_trace=False)
acc["c"] = 0.0; acc["n"] = 0
torch.cuda.synchronize(); t0 = time.perf_counter()
R = 50
for r in range(R):
    out = cbess.fit(None, y, 1, w, True, 1, 1, 20, 2, 1, True, 1, True, 10, seq, 1, 20, True, 5000, x_device_ptr=X.data_ptr(), n=n, p=p, want_trace=False)
torch.cuda.synchronize(); tot = time.perf_counter() - t0
print("per call: total %.3f ms, inside the C call %.3f ms, python wrapper %.3f ms; host phases inside C: %s sum %.3f" % (
    tot / R * 1e3, acc["c"] / R * 1e3, (tot - acc["c"]) / R * 1e3, {k2: round(v, 3) for k2, v in out["stats"]["host_ms"].items()},
    sum(out["stats"]["host_ms"].values())))
