"""Synthetic designs shaped like the reference's ``gen.data`` (R/R/gen.data.R:93-166, cortype 1, rho = 0;
python twin python/bess/gen_data.py:22-102).  R's RNG streams cannot be reproduced without R, so draws come
from ``numpy.random.Generator(PCG64(seed))``; the *recipe* (effect sizes, noise level, link, censoring) is
the reference's.  For cox the rows are returned already time-sorted with ``y = status`` -- what the
front-ends hand to ``bessCpp`` (R/R/bess.R:527-534, python/bess/linear.py:257-263)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SynthData:
    x: np.ndarray        # n x p, C-contiguous float64
    y: np.ndarray        # n (cox: status after time-sorting)
    beta: np.ndarray     # true coefficients
    time: np.ndarray | None = None


def gen_data(n, p, family="gaussian", k=10, seed=1, snr=10.0, scal=10.0, c=10.0, x=None) -> SynthData:
    rng = np.random.Generator(np.random.PCG64(seed))
    if x is None:
        x = rng.standard_normal((n, p))
    nonzero = rng.choice(p, size=k, replace=False)
    tbeta = np.zeros(p)
    m = 5.0 * np.sqrt(2.0 * np.log(p) / n)
    if family == "gaussian":
        tbeta[nonzero] = rng.uniform(m, 100.0 * m, k)
        sigma = np.sqrt((tbeta @ tbeta) / snr)
        y = x[:, nonzero] @ tbeta[nonzero] + rng.normal(0.0, sigma, n)
        return SynthData(x, y, tbeta)
    if family == "binomial":
        tbeta[nonzero] = rng.uniform(2 * m, 10 * m, k)
        sigma = np.sqrt((tbeta @ tbeta) / snr)
        eta = x[:, nonzero] @ tbeta[nonzero] + rng.normal(0.0, sigma, n)
        eta = np.clip(eta, -30, 30)
        pr = np.exp(eta) / (1 + np.exp(eta))
        y = rng.binomial(1, pr).astype(np.float64)
        return SynthData(x, y, tbeta)
    if family == "poisson":
        x = x / 16.0  # not in place: a caller-supplied design stays untouched
        tbeta[nonzero] = rng.uniform(2 * m, 10 * m, k)
        sigma = np.sqrt((tbeta @ tbeta) / snr)
        eta = np.clip(x[:, nonzero] @ tbeta[nonzero] + rng.normal(0.0, sigma, n), -30, 30)
        y = rng.poisson(np.exp(eta)).astype(np.float64)
        return SynthData(x, y, tbeta)
    if family == "cox":
        tbeta[nonzero] = rng.uniform(2 * m, 10 * m, k)
        time = (-np.log(rng.uniform(size=n)) / np.exp(x[:, nonzero] @ tbeta[nonzero])) ** (1.0 / scal)
        ctime = c * rng.uniform(size=n)
        status = (time < ctime).astype(np.float64)
        time = np.minimum(time, ctime)
        order = np.argsort(time, kind="stable")
        return SynthData(np.ascontiguousarray(x[order]), status[order], tbeta, time[order])
    raise ValueError("family should be 'gaussian', 'binomial', 'poisson' or 'cox'")


class data:
    """Result record of the reference's ``gen_data`` (python/bess/gen_data.py:4-8)."""

    def __init__(self, x, y, beta):
        self.x = x
        self.y = y
        self.beta = beta


def gen_data_reference(n, p, family, k, rho=0, sigma=1, beta=None, censoring=True, c=1, scal=10):
    """The reference's Python generator, argument for argument (python/bess/gen_data.py:22-102): the banded design
    x = X + rho * (left shift + right shift) of the centred, sqrt(n)-normalised iid normal X (``gen.data`` cortype 3,
    R/R/gen.data.R:167-245), draws from numpy's GLOBAL random state like the reference, ``y`` of a cox problem is the
    unsorted n x 2 array [time, status] that ``PdasCox.fit`` sorts itself (linear.py:257-263).  Served as
    ``bess.gen_data.gen_data`` by ``bess_b200.compat.install_as_bess``; ``gen_data`` above (seeded PCG64 streams,
    cortype 1, rows pre-sorted for cox) is what the parity tests and the bench use."""
    X = np.random.normal(0, 1, n * p).reshape(n, p)
    X = X - X.mean(axis=0, keepdims=True)
    X = np.sqrt(n) * X / np.sqrt((X ** 2).sum(axis=0, keepdims=True))
    zero = np.zeros((n, 1))
    x = X + rho * (np.hstack((zero, X[:, 0:(p - 2)], zero)) + np.hstack((zero, X[:, 2:p], zero)))
    full, nonzero = np.arange(p), np.zeros(k, int)
    for i in range(k):  # sampling without replacement, one draw at a time (gen_data.py:11-19)
        z = np.random.choice(full, 1)
        nonzero[i] = z[0]
        full = np.delete(full, np.where(full == z))
    tbeta = np.zeros(p)
    m = 5 * (1 if family == "gaussian" else sigma) * np.sqrt(2 * np.log(p) / n)
    if beta is None:
        tbeta[nonzero] = np.random.uniform(m, 100 * m, k) if family == "gaussian" else np.random.uniform(2 * m, 10 * m, k)
    else:
        tbeta = beta
    if family == "gaussian":
        return data(x, np.matmul(x, tbeta) + sigma * np.random.normal(0, 1, n), tbeta)
    if family == "binomial":
        xb = np.clip(np.matmul(x, tbeta), -30, 30)
        return data(x, np.random.binomial(1, np.exp(xb) / (1 + np.exp(xb))), tbeta)
    if family == "poisson":
        x = x / 16
        xb = np.clip(np.matmul(x, tbeta), -30, 30)
        return data(x, np.random.poisson(lam=np.exp(xb)), tbeta)
    if family == "cox":
        time = np.power(-np.log(np.random.uniform(0, 1, n)) / np.exp(np.matmul(x, tbeta)), 1 / scal)
        if censoring:
            ctime = c * np.random.uniform(0, 1, n)
            status = (time < ctime) * 1
            print("censoring rate:" + str(1 - sum(status) / n))
            time = np.minimum(time, ctime)
        else:
            status = np.ones(n)
            print("no censoring")
        return data(x, np.hstack((time.reshape((-1, 1)), status.reshape((-1, 1)))), tbeta)
    raise ValueError("Family should be 'gaussian', 'binomial', 'possion', or 'cox'")


def gen_response_device(X, family, k, seed=1, snr=10.0, scal=10.0, c=10.0):
    """The y of ``gen.data`` for a design that lives in HBM (a ``torch`` tensor from ``gen_design_device``): only the k
    active columns are touched -- gathered and multiplied on the device (``X[:, nonzero] @ beta``), n numbers come back --,
    the draws of the noise / the response (n numbers) are numpy's.  Same recipe as ``gen_data`` above; for cox the caller
    gets (time, status) and must sort the rows of X by time itself (a 4 GB row permutation is the caller's call).
    Returns (y, beta_true, nonzero) and for cox ((time, status), beta_true, nonzero)."""
    import torch
    n, p = X.shape
    rng = np.random.Generator(np.random.PCG64(seed))
    nonzero = np.sort(rng.choice(p, size=k, replace=False))
    m = 5.0 * np.sqrt(2.0 * np.log(p) / n)
    tb = rng.uniform(m, 100.0 * m, k) if family == "gaussian" else rng.uniform(2 * m, 10 * m, k)
    eta = (X[:, torch.as_tensor(nonzero, device=X.device)] @ torch.as_tensor(tb, device=X.device)).cpu().numpy()
    sigma = np.sqrt((tb @ tb) / snr)
    if family == "gaussian":
        return eta + rng.normal(0.0, sigma, n), tb, nonzero
    # binomial / poisson: the linear predictor carries no noise term (python/bess/gen_data.py:44-73); the caller scales a
    # poisson design by 1/16 first, as the reference does
    if family == "binomial":
        e = np.clip(eta, -30, 30)
        return rng.binomial(1, np.exp(e) / (1 + np.exp(e))).astype(np.float64), tb, nonzero
    if family == "poisson":
        e = np.clip(eta, -30, 30)
        return rng.poisson(np.exp(e)).astype(np.float64), tb, nonzero
    if family == "cox":
        time = (-np.log(rng.uniform(size=n)) / np.exp(eta)) ** (1.0 / scal)
        ctime = c * rng.uniform(size=n)
        return (np.minimum(time, ctime), (time < ctime).astype(np.float64)), tb, nonzero
    raise ValueError("family should be 'gaussian', 'binomial', 'poisson' or 'cox'")


def gen_design_device(n, p, rho=0.0, seed=1, device=0, cortype=1):
    """The x of ``gen.data`` (R/R/gen.data.R:110-118, cortype 1: rows ~ MVN(0, Sigma), Sigma_jk = rho^|j-k|) drawn on the
    GPU by the library's own kernel (``bess_b200_gen_design``): a row-major n x p fp64 ``torch`` tensor in HBM that can be
    handed to ``cbess.fit(..., x_device_ptr=X.data_ptr())`` -- torch only owns the memory."""
    import torch
    from . import _lib
    lib = _lib.load()
    _lib.require_gpu()
    X = torch.empty((n, p), dtype=torch.float64, device=f"cuda:{device}")
    # cortype 1: Sigma_jk = rho^|j-k|; 2: exchangeable rho + (1 - rho) I; 3: the banded design (gen.data cortype 3, the
    # Python generator of the reference)
    _lib.check(lib.bess_b200_gen_design_cortype(X.data_ptr(), int(n), int(p), int(p), float(rho), int(seed), int(cortype),
                                                int(device)))
    return X
