"""Size-independent properties at BASELINE.json's full sizes (-m gpu), through the C-ABI: symmetries the model has and
the kernels' tilings, reductions and fold geometry must not break.  (The same sizes are also compared with the CPU
oracle directly in tests/test_gpu_round2.py; these properties need no checker and run in milliseconds.)"""
import numpy as np
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from gpscore_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("score", ["crps", "nlml"])
def test_full_gp_n10000_row_permutation_invariance(ctx, score):
    from gpscore_b200 import synth
    X, y = synth.kin40k_like(10000)
    theta = synth.hyper_point("P1")
    p = np.random.default_rng(1).permutation(10000)
    ctx.set_data(_dev(X), _dev(y))
    v0, g0 = ctx.full_eval(theta, score)
    ctx.set_data(_dev(X[p]), _dev(y[p]))
    v1, g1 = ctx.full_eval(theta, score)
    assert abs(v0 - v1) <= 1e-10 * abs(v0)
    assert relerr(g1, g0) <= 1e-8


def test_fitc20_n1e6_inducing_permutation_and_row_block_additivity(ctx):
    """Permuting the inducing points permutes grad_U and changes nothing else; and the NLML's data-fit part is additive
    over row blocks only through the shared M x M matrices, so two half-size problems do NOT add up — but the two
    orderings [A; B] and [B; A] of the same rows must agree (tiles, tickets and reductions see different rows)."""
    from gpscore_b200 import synth
    n = 1000000
    X, y = synth.kin40k_like(n, seed=7)
    theta = synth.hyper_point("P1")
    U = synth.inducing_init(20)
    q = np.random.default_rng(2).permutation(20)
    ctx.set_data(_dev(X), _dev(y))
    v0, g0, gu0 = ctx.fitc_eval(theta, U, "crps")
    v1, g1, gu1 = ctx.fitc_eval(theta, U[q], "crps")
    assert abs(v0 - v1) <= 1e-11 * abs(v0)
    assert relerr(g1, g0) <= 1e-8 and relerr(gu1, gu0[q]) <= 1e-8
    h = n // 2 + 12345
    Xs, ys = np.concatenate([X[h:], X[:h]]), np.concatenate([y[h:], y[:h]])
    ctx.set_data(_dev(Xs), _dev(ys))
    v2, g2, gu2 = ctx.fitc_eval(theta, U, "crps")
    assert abs(v0 - v2) <= 1e-11 * abs(v0)
    assert relerr(g2, g0) <= 1e-8 and relerr(gu2, gu0) <= 1e-8


@pytest.mark.parametrize("kind", ["dss", "kc"])
@pytest.mark.parametrize("m_ind", [20, 64])
def test_fitc_block_objectives_n10000_depend_on_folds_as_sets(ctx, kind, m_ind):
    """Rows permuted inside each quarter leave the 4-fold objectives unchanged (row kernels at M = 20, matrix form at
    M = 64: fold-aligned tiles / masked 16-rounded column ranges must not leak rows across folds)."""
    from gpscore_b200 import synth
    n = 10000
    X, y = synth.kin40k_like(n)
    theta = synth.hyper_point("P1")
    rng = np.random.default_rng(3)
    U = X[rng.choice(n, m_ind, replace=False)] + 0.01 * rng.standard_normal((m_ind, 8))
    p = np.concatenate([f * 2500 + rng.permutation(2500) for f in range(4)])
    ctx.set_data(_dev(X), _dev(y))
    v0, g0, gu0 = ctx.fitc_eval(theta, U, kind)
    ctx.set_data(_dev(X[p]), _dev(y[p]))
    v1, g1, gu1 = ctx.fitc_eval(theta, U, kind)
    assert abs(v0 - v1) <= 1e-10 * abs(v0)
    assert relerr(g1, g0) <= 1e-7 and relerr(gu1, gu0) <= 1e-7
    r = rng.permutation(n)
    ctx.set_data(_dev(X[r]), _dev(y[r]))
    v2 = ctx.fitc_eval(theta, U, kind)[0]
    assert abs(v2 - v0) > 1e-8 * abs(v0)          # other folds, another objective
