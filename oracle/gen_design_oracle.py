"""TEST INFRASTRUCTURE ONLY -- numpy restatement of bess_b200/csrc/gen_design.cu, the device-side design generator
(the x of gen.data, /root/reference/R/R/gen.data.R:110-118, cortype 1: rows ~ MVN(0, Sigma), Sigma_jk = rho^|j-k|).

The reference draws x with R's mvrnorm (Mersenne-Twister + inversion); that stream cannot be reproduced without R
(SURVEY 8d), so parity here is (a) the published Philox4x32-10 algorithm against its known-answer vectors (Random123
kat_vectors), (b) the device stream against this restatement, (c) the distribution (moments, lag correlations)."""
import math

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3" (SC'11).  Vectorised over the counters."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK for v in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def normal_at(seed, i, j):
    """z_ij of gen_design.cu: counter (j as signed 64-bit, i, 0), key = seed; Box-Muller, cosine branch."""
    i = np.asarray(i, dtype=np.int64)
    j = np.asarray(j, dtype=np.int64)
    ju = j.astype(np.uint64)  # two's complement
    r0, r1, r2, r3 = philox4x32_10(ju & MASK, ju >> np.uint64(32), i.astype(np.uint64), np.zeros_like(ju),
                                   seed & 0xFFFFFFFF, seed >> 32)
    a = (r0 << np.uint64(32)) | r1
    b = (r2 << np.uint64(32)) | r3
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 0.5) * 2.0 ** -53
    u2 = ((b >> np.uint64(11)).astype(np.float64) + 0.5) * 2.0 ** -53
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def design(n, p, rho=0.0, seed=1):
    """n x p design: the stationary AR(1) recurrence along the columns, started far enough to the left of column 0
    that the missing tail is below 2^-60 (the device starts every segment its own warm-up; both agree to ~1e-15)."""
    warm = 0 if rho == 0.0 else int(math.ceil(60.0 * math.log(2.0) / -math.log(abs(rho)))) + 64
    jj = np.arange(-warm, p, dtype=np.int64)
    ii = np.arange(n, dtype=np.int64)
    z = normal_at(seed, ii[:, None], jj[None, :])
    if rho == 0.0:
        return z
    s = math.sqrt(1.0 - rho * rho)
    x = np.empty_like(z)
    prev = np.zeros(n)
    for t in range(z.shape[1]):
        prev = rho * prev + s * z[:, t]
        x[:, t] = prev
    return x[:, warm:]


def design_cortype(n, p, rho, seed, cortype):
    """gen_design.cu launch_gen_design_cortype: 1 = design(); 2 = exchangeable (R/R/gen.data.R:114-116), common factor
    z_i,-1; 3 = the banded design (gen.data.R:167-181, python/bess/gen_data.py:25-30) on the iid stream."""
    if cortype == 1:
        return design(n, p, rho, seed)
    ii = np.arange(n, dtype=np.int64)[:, None]
    if cortype == 2:
        f = normal_at(seed, ii, np.full((1, 1), -1, dtype=np.int64))
        return math.sqrt(rho) * f + math.sqrt(1.0 - rho) * normal_at(seed, ii, np.arange(p, dtype=np.int64)[None, :])
    X = design(n, p, 0.0, seed)
    X = X - X.mean(axis=0, keepdims=True)
    X = math.sqrt(n) * X / np.sqrt((X ** 2).sum(axis=0, keepdims=True))
    zero = np.zeros((n, 1))
    return X + rho * (np.hstack((zero, X[:, 0:(p - 2)], zero)) + np.hstack((zero, X[:, 2:p], zero)))
