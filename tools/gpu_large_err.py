import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pdas_oracle as orc
from bess_b200.engine import GpuEngine
from bess_b200.gen_data import gen_data
from tests.helpers import rel_err
FAM = {"gaussian": (1, 1), "binomial": (2, 2), "poisson": (3, 2), "cox": (4, 3)}
import tests.test_gpu_parity as T
import inspect
src = inspect.getsource(T)
import re
m = re.search(r'@pytest.mark.parametrize\("fam,n,p,T", (\[.*?\])\)\ndef test_large_support', src, re.S)
cases = eval(m.group(1))
for fam, n, p, Tq in cases:
    model_type, data_type = FAM[fam]
    d = gen_data(n, p, fam, 10, seed=31)
    w = np.ones(n)
    eng = GpuEngine()
    eng.load(d.x, d.y, w, model_type)
    eng.normalize(data_type, True)
    eng.setup_chains(0, None, Tq, 20, True)
    data = orc.make_data(d.x, d.y, w, data_type, True, model_type)
    st = orc.PathState(data, model_type, 3, False, 0, None, 20, True)
    r = eng.run_batch(Tq, [0], True)
    o = orc.pdas_fit(data, model_type, Tq, np.zeros(p), 0.0, st.full_mask, st.xtx_full, 20)
    XA = data.x[:, o.A]
    c = np.linalg.cond(XA.T @ XA)
    print(fam, n, p, Tq, "same A", r["A"][0].tolist() == o.A.tolist(), "rel_err", rel_err(r["bA"][0], o.beta[o.A]), "cond(X_A'X_A)", c, flush=True)
    eng.close()
