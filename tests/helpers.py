"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-8  # north_star: coefficients and losses within 1e-8 relative in fp64


def golden_names():
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    m = g["meta"]
    g["model_type"], g["data_type"], g["path_type"], g["is_cv"], g["K"], g["ic_type"], g["smax"], g["scr"] = (
        int(m[0]), int(m[1]), int(m[2]), bool(m[3]), int(m[4]), int(m[5]), int(m[6]), int(m[7]))
    return g


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b)) / scale) if a.size else 0.0


def assert_same_support(beta_a, beta_b):
    sa = np.nonzero(beta_a)[0]
    sb = np.nonzero(beta_b)[0]
    assert sa.tolist() == sb.tolist(), f"support differs: {sa.tolist()} vs {sb.tolist()}"
