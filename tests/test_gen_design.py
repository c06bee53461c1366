"""Device-side gen.data design generator (bess_b200/csrc/gen_design.cu; R/R/gen.data.R:110-118, cortype 1).
CPU: the numpy restatement against the published Philox4x32-10 known-answer vectors and the target distribution.
GPU: the device stream against the restatement, reproducibility, and the distribution at a size that matters."""
import numpy as np
import pytest

from oracle import gen_design_oracle as gd


def test_philox_known_answer_vectors():
    """Random123 kat_vectors, philox4x32 10 rounds: counter words, key words -> output words."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        got = gd.philox4x32_10(*[np.array([c]) for c in ctr], *key)
        assert tuple(int(v[0]) for v in got) == exp


@pytest.mark.parametrize("rho", [0.0, 0.6, -0.3])
def test_oracle_design_has_the_gen_data_covariance(rho):
    x = gd.design(300, 4000, rho, seed=11)
    assert abs(x.mean()) < 5e-3 and abs(x.std() - 1.0) < 5e-3
    for lag in (1, 2, 3):
        c = np.corrcoef(x[:, :-lag].ravel(), x[:, lag:].ravel())[0, 1]
        assert abs(c - rho ** lag) < 6e-3  # Sigma_jk = rho^|j-k|
    assert abs(np.corrcoef(x[:-1].ravel(), x[1:].ravel())[0, 1]) < 6e-3  # rows independent
    assert np.array_equal(x, gd.design(300, 4000, rho, seed=11)) and not np.array_equal(x, gd.design(300, 4000, rho, seed=12))


@pytest.mark.gpu
@pytest.mark.parametrize("n,p,rho", [(7, 33, 0.0), (64, 5000, 0.0), (37, 9001, 0.5), (16, 20000, 0.9), (5, 100, -0.4)])
def test_device_stream_matches_the_restatement(n, p, rho):
    from bess_b200.gen_data import gen_design_device
    x = gen_design_device(n, p, rho, seed=2024).cpu().numpy()
    ref = gd.design(n, p, rho, seed=2024)
    assert x.shape == (n, p)
    # log / cos differ by an ulp or two between the CUDA math library and libm; the streams are the same
    assert np.max(np.abs(x - ref)) < 1e-12
    assert np.array_equal(x, gen_design_device(n, p, rho, seed=2024).cpu().numpy())


@pytest.mark.gpu
def test_device_design_distribution_and_use():
    """p = 200000 columns, rho = 0.7: moments and lag correlations; the generated design then goes through a fit without
    ever touching the host (x_device_ptr)."""
    from bess_b200 import cbess
    from bess_b200.gen_data import gen_design_device
    n, p, rho = 500, 200000, 0.7
    X = gen_design_device(n, p, rho, seed=5)
    x = X[:, :50000].cpu().numpy()
    assert abs(x.mean()) < 2e-3 and abs(x.std() - 1.0) < 2e-3
    for lag in (1, 2):
        assert abs(np.corrcoef(x[:, :-lag].ravel(), x[:, lag:].ravel())[0, 1] - rho ** lag) < 3e-3
    # segment seams (the kernel starts each 4096-column segment with its own warm-up): correlation across a seam
    seam = np.corrcoef(x[:, 4095], x[:, 4096])[0, 1]
    assert abs(seam - rho) < 0.08
    nz = np.array([10, 5000, 123456])
    y = (X[:, nz].cpu().numpy() @ np.array([3.0, -2.0, 4.0])) + np.random.default_rng(0).normal(0, 0.5, n)
    out = cbess.fit(None, y, 1, np.ones(n), True, 1, 1, 20, 2, 1, True, 3, False, 5, [1, 2, 3, 4], 1, 4, True, 2000,
                    x_device_ptr=X.data_ptr(), n=n, p=p, want_trace=False)
    assert set(nz.tolist()) <= set(np.nonzero(out["beta"])[0].tolist())


@pytest.mark.parametrize("cortype,rho", [(2, 0.4), (3, 0.5)])
def test_oracle_cortype_2_and_3_have_the_gen_data_structure(cortype, rho):
    """cortype 2 (R/R/gen.data.R:114-116): every pair of columns correlates rho.  cortype 3 (gen.data.R:167-181,
    python/bess/gen_data.py:25-30): neighbours 2 rho / (1 + 2 rho^2), second neighbours rho^2 / (1 + 2 rho^2), none beyond."""
    x = gd.design_cortype(400, 1500, rho, 7, cortype)
    c = np.corrcoef(x, rowvar=False)
    if cortype == 2:
        off = c[np.triu_indices(1500, 1)]
        assert abs(off.mean() - rho) < 0.05 and abs(x.std() - 1.0) < 0.05  # 400 draws of the common factor
    else:
        d1, d2, d3 = np.diag(c, 1)[1:-1], np.diag(c, 2)[1:-1], np.diag(c, 3)
        assert abs(d1.mean() - 2 * rho / (1 + 2 * rho ** 2)) < 0.02
        assert abs(d2.mean() - rho ** 2 / (1 + 2 * rho ** 2)) < 0.02 and abs(d3.mean()) < 0.02
        assert np.allclose(x[:, 0].std(), 1.0, atol=1e-9)  # the edge columns are the normalised X itself


@pytest.mark.gpu
@pytest.mark.parametrize("n,p,rho,cortype", [(40, 3000, 0.3, 2), (33, 2500, 0.5, 3), (17, 5, 0.25, 3), (64, 1025, 0.5, 3)])
def test_device_cortype_2_and_3_match_the_restatement(n, p, rho, cortype):
    from bess_b200.gen_data import gen_design_device
    x = gen_design_device(n, p, rho, seed=99, cortype=cortype).cpu().numpy()
    ref = gd.design_cortype(n, p, rho, 99, cortype)
    assert np.max(np.abs(x - ref)) < 1e-10


@pytest.mark.gpu
def test_device_response_recipe_feeds_a_fit():
    """gen_response_device: the y of gen.data from a design that never leaves HBM (only the k active columns are read)."""
    from bess_b200 import cbess
    from bess_b200.gen_data import gen_design_device, gen_response_device
    n, p = 400, 60000
    X = gen_design_device(n, p, 0.5, seed=3, cortype=3)
    for fam, (mt, dt) in {"gaussian": (1, 1), "binomial": (2, 2), "poisson": (3, 2)}.items():
        Xf = X / 16.0 if fam == "poisson" else X
        y, tb, nz = gen_response_device(Xf, fam, 5, seed=17)
        out = cbess.fit(None, y, dt, np.ones(n), True, 1, mt, 20, 2, 1, True, 3, False, 5, [1, 2, 3, 4, 5, 6], 1, 6, True, 3000,
                        x_device_ptr=Xf.data_ptr(), n=n, p=p, want_trace=False)
        got = set(np.nonzero(out["beta"])[0].tolist())
        assert len(got & set(nz.tolist())) >= 3, (fam, sorted(got), nz.tolist())
