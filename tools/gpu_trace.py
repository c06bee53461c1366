"""Time line of one resident-path launch (BESS_B200_TRACE): per round, who is the last owner to finish and how long the
sweeper steps / hops take."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BESS_B200_TRACE"] = "/tmp/lp_trace.txt"
from bess_b200 import cbess
n, p, k = 1000, 500000, 10
g = torch.Generator(device="cuda").manual_seed(5)
X = torch.randn(n, p, dtype=torch.float64, device="cuda", generator=g)
rng = np.random.default_rng(5)
nz = np.sort(rng.choice(p, k, replace=False))
beta = rng.uniform(1, 5, k)
y = (X[:, torch.as_tensor(nz, device="cuda")] @ torch.as_tensor(beta, device="cuda")).cpu().numpy() + rng.normal(0, 3, n)
w = np.ones(n); seq = np.arange(1, 21)
for r in range(4):
    out = cbess.fit(None, y, 1, w, True, 1, 1, 20, 2, 1, True, 1, True, 10, seq, 1, 20, True, 5000, x_device_ptr=X.data_ptr(), n=n, p=p, want_trace=False)
rows = np.loadtxt("/tmp/lp_trace.txt", dtype=np.int64)
t0 = rows[:, 2][rows[:, 2] > 0].min()
own = {}
swp = {}
for c, q, a, b in rows:
    (swp if c == 32 else own.setdefault(int(c), {}))[int(q)] = ((a - t0) / 1e3, (b - t0) / 1e3)
print("sweeper 0 steps (start, end, dur us):")
for q in sorted(swp)[:24]:
    a, b = swp[q]
    print(f"  step {q:2d} g{q % 2}: {a:8.1f} -> {b:8.1f}  ({b - a:5.1f})")
print("owner phases by round (group 0 = even positions): start spread, durations, end of slowest")
for it in range(1, 14):
    for g in (0, 1):
        cs = [c for c in sorted(own) if c % 2 == g and it in own[c]]
        if not cs:
            continue
        st = [own[c][it][0] for c in cs]
        en = [own[c][it][1] for c in cs]
        print(f"  round {it:2d} g{g}: start {min(st):8.1f}  durs {[round(e - s, 1) for s, e in zip(st, en)]}  last end {max(en):8.1f}")
