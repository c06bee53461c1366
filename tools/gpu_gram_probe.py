"""Probe of the chain kernels' weighted Gram (one CTA, rows of one cluster slice) against numpy, timed with clock64."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bess_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    _lib.require_gpu()
    lib.bess_b200_debug_gram.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                         C.c_void_p]
    rng = np.random.default_rng(0)
    cases = [(225, 203), (225, 103), (225, 165), (225, 265), (113, 203), (450, 203), (250, 32), (250, 48), (250, 57), (250, 64), (250, 80), (250, 97), (37, 130), (1800, 42), (1800, 64)]
    if os.environ.get("GRAM_CASES"):
        cases = [tuple(int(v) for v in c.split("x")) for c in os.environ["GRAM_CASES"].split(",")]
    reps = int(os.environ.get("GRAM_REPS", "3"))
    impls = [int(v) for v in os.environ.get("GRAM_IMPLS", "1,2").split(",")]
    for nrows, mm in cases:
        ldv = (mm + 1) & ~1
        V = np.zeros((nrows, ldv))
        V[:, :mm] = rng.standard_normal((nrows, mm))
        w = rng.uniform(0.001, 0.25, nrows)
        ref = (V[:, :mm] * w[:, None]).T @ V[:, :mm]
        line = f"rows={nrows:4d} m={mm:4d}"
        for impl in impls:
            S = np.zeros((ldv, ldv))
            t = C.c_double(0)
            rc = lib.bess_b200_debug_gram(V.ctypes.data, ldv, nrows, mm, w.ctypes.data, S.ctypes.data, impl, reps, C.byref(t))
            if rc != 0:
                line += f"   impl{impl}: rc={rc} {lib.bess_b200_last_error().decode()}"
                continue
            err = np.max(np.abs(S[:mm, :mm] - ref)) / np.max(np.abs(ref))
            fma = nrows * mm * (mm + 1) / 2
            line += f"   {'fma ' if impl == 1 else 'dmma'}: {t.value / 1965:7.1f} us err {err:.1e} ({fma / t.value:5.1f} useful FMA/clk)"
        print(line, flush=True)


if __name__ == "__main__":
    main()
