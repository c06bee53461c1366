"""Probe of the chain kernels' normal-equation solvers: random SPD bordered systems through one CTA of the probe kernel
(bess_b200_debug_solve), checked against numpy and timed with clock64.  impl 0 = panel-major (DMMA trailing update),
impl 1 = the round-1 packed in-smem Cholesky."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bess_b200 import _lib  # noqa: E402


IMPLS = [int(a) for a in os.environ.get("IMPLS", "0,1").split(",")]


def main():
    lib = _lib.load()
    _lib.require_gpu()
    lib.bess_b200_debug_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(0)
    sizes = [int(a) for a in sys.argv[1:]] or [5, 16, 17, 65, 96, 100, 128, 163, 201, 202, 224, 230, 240, 264, 300, 400, 500]
    for mm in sizes:
        n = 4 * mm + 50
        V = rng.standard_normal((n, mm))
        w = rng.uniform(0.001, 0.25, n)
        z = rng.standard_normal(n)
        A = (V * w[:, None]).T @ V
        rhs = (V * w[:, None]).T @ z
        lds = (mm + 3) & ~1
        S = np.zeros((mm + 1, lds))
        S[:mm, :mm] = A
        S[mm, :mm] = rhs
        ref = np.linalg.solve(A, rhs)
        line = f"mm={mm:4d}"
        for impl in IMPLS:
            if impl == 1 and (mm * (mm + 3) // 2 + 600 > 28000):
                line += "   packed: n/a"
                continue
            x = np.zeros(mm)
            t = C.c_double(0)
            lib.bess_b200_debug_set(2, 1)
            rc = lib.bess_b200_debug_solve(S.ctypes.data, lds, mm, x.ctypes.data, impl, 3, C.byref(t))
            buf = (C.c_ulonglong * 32)()
            lib.bess_b200_debug_get(buf)
            lib.bess_b200_debug_set(2, 0)
            ph = " ".join(f"{nm}={buf[i] / 1965:.1f}" for i, nm in ((10, "load"), (13, "rows"), (14, "trail"), (11, "stage"), (12, "back")) if buf[i])
            if rc != 0:
                line += f"   impl{impl}: rc={rc} {lib.bess_b200_last_error().decode()}"
                continue
            err = np.max(np.abs(x - ref)) / np.max(np.abs(ref))
            line += f"   { {0: 'panel', 1: 'packed'}.get(impl, 'variant%d' % impl)}: {t.value / 1965:7.1f} us err {err:.1e} [{ph}]"
        print(line, flush=True)


if __name__ == "__main__":
    main()
