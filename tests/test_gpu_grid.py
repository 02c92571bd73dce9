"""Grid-sweep parity (-m gpu): gps_grid_eval against the numpy restatement of the R functions of
contour-plot.R (CP:43-85) on the script's own kind of data (n = 20, x = seq(-6, 6), CP:33-39),
for all four objectives; plus the n > 128 route through the blocked factorisation."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from gpscore_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def _toy(n, seed=0):
    from oracle import gp_oracle as O
    rng = np.random.default_rng(seed)
    x = np.linspace(-6, 6, n)                                   # CP:35
    K = O.rbf_R(x, x, l=1.0, k=1.0) + 1e-10 * np.eye(n)
    y = np.linalg.cholesky(K) @ rng.standard_normal(n) + 0.1 * rng.standard_normal(n)   # CP:36-38
    return x, y.reshape(-1, 1)


ORACLE = {"nlml": "cal_NLML", "crps": "cal_m_crps", "wrong_crps": "wrong_cal_m_crps", "logs": "cal_m_logs"}


@pytest.mark.parametrize("which", ["nlml", "crps", "wrong_crps", "logs"])
@pytest.mark.parametrize("n", [20, 128])
def test_grid_small_vs_r_twins(ctx, which, n):
    from oracle import gp_oracle as O
    x, y = _toy(n)
    l_range = np.linspace(0.05, 2, 12)                          # CP:88 range (0.01 is singular at n = 128)
    noise_range = np.linspace(0.05, 1, 10)                      # CP:109 range
    Lg, Sg = np.meshgrid(l_range, noise_range, indexing="ij")
    got = ctx.grid_eval(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), Lg.ravel(), Sg.ravel(), which)
    fn = getattr(O, ORACLE[which])
    want = np.array([fn(x, y, l, s) for l, s in zip(Lg.ravel(), Sg.ravel())])
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) <= 1e-8
    from gpscore_b200 import dist as D
    assert D.grid_matrix(got, 12, 10).shape == (10, 12)         # CP:114 layout: rows = noise s.d.


def test_grid_reference_ranges_50x50(ctx):
    """The full 50 x 50 sweep of CP:88/109 at n = 20, CRPS: spot-checked at 60 random points."""
    from oracle import gp_oracle as O
    x, y = _toy(20, seed=1)
    Lg, Sg = np.meshgrid(np.linspace(0.01, 2, 50), np.linspace(0.01, 1, 50), indexing="ij")
    got = ctx.grid_eval(x, y, Lg.ravel(), Sg.ravel(), "crps")   # host inputs
    idx = np.random.default_rng(0).choice(2500, 60, replace=False)
    want = np.array([O.cal_m_crps(x, y, Lg.ravel()[i], Sg.ravel()[i]) for i in idx])
    assert np.max(np.abs(got[idx] - want) / np.maximum(np.abs(want), 1.0)) <= 1e-8


@pytest.mark.parametrize("which", ["nlml", "crps", "wrong_crps", "logs"])
def test_grid_large_n(ctx, which):
    from oracle import gp_oracle as O
    n = 300
    rng = np.random.default_rng(2)
    x = np.sort(rng.uniform(-6, 6, n))
    y = (np.sin(x) + 0.2 * rng.standard_normal(n)).reshape(-1, 1)
    ls = np.array([0.3, 0.8, 1.5])
    sd = np.array([0.2, 0.5, 0.9])
    got = ctx.grid_eval(x, y, ls, sd, which)
    fn = getattr(O, ORACLE[which])
    want = np.array([fn(x, y, l, s) for l, s in zip(ls, sd)])
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) <= 1e-8
