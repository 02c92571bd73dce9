"""Edge cases (-m gpu): tiny, ragged and degenerate sizes through the C-ABI, against the oracle."""
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


@pytest.mark.parametrize("n,d", [(2, 1), (3, 5), (127, 2), (128, 1), (129, 8), (257, 3)])
def test_full_tiny_and_tile_boundaries(ctx, n, d):
    from oracle import gp_oracle as O
    rng = np.random.default_rng(n * 10 + d)
    X = rng.standard_normal((n, d))
    y = rng.standard_normal((n, 1))
    theta = np.concatenate([[0.3], 0.2 * rng.standard_normal(d), [-0.7]])
    ctx.set_data(_dev(X), _dev(y))
    for score in ("crps", "logs", "nlml"):
        val, grad = ctx.full_eval(theta, score)
        oval, ograd = O.full_obj_grad(X, y, theta, O.SCORES[score])
        assert abs(val - oval) <= 1e-8 * max(abs(oval), 1e-12), (score, val, oval)
        assert relerr(grad, ograd) <= 1e-6, score
    Xs = rng.standard_normal((1, d))                      # a single test row
    mean, var = ctx.full_predict(theta, _dev(Xs))
    om, ov = O.full_predict(X, y, Xs, theta)
    assert relerr(mean.cpu().numpy(), om) <= 1e-8 and relerr(var.cpu().numpy(), ov) <= 1e-8


@pytest.mark.parametrize("n,m,d", [(4, 1, 1), (5, 3, 2), (130, 8, 1), (300, 17, 15)])
def test_fitc_tiny_and_max_dims(ctx, n, m, d):
    from oracle import gp_oracle as O
    rng = np.random.default_rng(n + m + d)
    X = rng.standard_normal((n, d))
    y = rng.standard_normal((n, 1))
    U = rng.standard_normal((m, d))
    theta = np.concatenate([[0.1], 0.3 * rng.standard_normal(d) + 0.5, [-0.5]])
    ctx.set_data(_dev(X), _dev(y))
    for score in ("crps", "logs", "nlml"):
        val, g, gU = ctx.fitc_eval(theta, U, score)
        oval, og, ogU, _ = O.fitc_obj_grad(X, y, U, theta, O.SCORES[score])
        assert abs(val - oval) <= 1e-8 * abs(oval), score
        assert relerr(g, og) <= 1e-6 and relerr(gU, ogU) <= 1e-6, score
    Xs = rng.standard_normal((3, d))
    mean, var = ctx.fitc_predict(theta, U, _dev(Xs))
    om, ov = O.fitc_predict(X, y, U, Xs, theta)
    assert relerr(mean.cpu().numpy(), om) <= 1e-8 and relerr(var.cpu().numpy(), ov) <= 1e-7


def test_argument_errors_are_reported(ctx):
    from gpscore_b200 import lib as L
    X = np.random.default_rng(0).standard_normal((50, 2))
    y = np.zeros((50, 1))
    ctx.set_data(_dev(X), _dev(y))
    with pytest.raises(L.GpsError):                       # no accumulator layout beyond the matrix form
        ctx.fitc_acc_len(4097)
    with pytest.raises(L.GpsError):                       # the block objectives' four folds need 4 | N (here N = 50)
        ctx.fitc_eval(np.zeros(4), np.random.default_rng(1).standard_normal((33, 2)), "kc")
    with pytest.raises(L.GpsError):                       # M beyond the matrix form
        ctx.fitc_eval(np.zeros(4), np.zeros((4097, 2)), "crps")
    with pytest.raises(ValueError):                       # theta of the wrong length
        ctx.full_eval(np.zeros(7), "crps")
    with pytest.raises(L.GpsError):                       # kc is a FITC objective
        ctx.full_eval(np.zeros(4), "kc")
    with pytest.raises(L.GpsError):                       # no LOO output after an NLML evaluation
        ctx.full_eval(np.zeros(4), "nlml")
        ctx.full_loo()


def test_full_size_prediction_is_finite_and_bounded(ctx):
    """BASELINE's prediction size: N = 10 000 train, T = 30 000 test rows (chunked by N rows)."""
    from gpscore_b200 import synth
    X, y, Xs, ys = synth.kin40k_like(10000, 30000)
    theta = synth.hyper_point("P2")
    ctx.set_data(_dev(X), _dev(y))
    mean, var = ctx.full_predict(theta, _dev(Xs))
    assert mean.shape == (30000, 1) and bool(torch.isfinite(mean).all()) and bool(torch.isfinite(var).all())
    sn2, sf2 = np.exp(theta[-1]), np.exp(theta[0])
    assert float(var.min()) >= sn2 * (1 - 1e-9) and float(var.max()) <= sn2 + sf2 + 1e-9
    m = ctx.test_metrics(mean, var, _dev(ys), _dev(y))
    assert 0.0 < m["smse"] < 0.5 and 0.8 < m["coverage"] <= 1.0
    # a chunk boundary is invisible: rows 9990..10010 predicted alone give the same numbers
    m2, v2 = ctx.full_predict(theta, _dev(Xs[9990:10010]))
    assert relerr(m2.cpu().numpy(), mean[9990:10010].cpu().numpy()) <= 1e-10
    assert relerr(v2.cpu().numpy(), var[9990:10010].cpu().numpy()) <= 1e-10


def test_two_devices_in_one_process():
    """Contexts on two GPUs of one process (kernel attributes are configured per device): both reproduce
    the single-device result.  Skipped on a one-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from gpscore_b200 import api, synth
    X, y = synth.kin40k_like(1500, seed=5)
    theta = synth.hyper_point("P1")
    U = synth.inducing_init(20, seed=6)
    outs = []
    for dev in (0, 1):
        c = api.Context(dev)
        c.set_data(torch.from_numpy(X).to("cuda:%d" % dev), torch.from_numpy(y).to("cuda:%d" % dev))
        outs.append((c.full_eval(theta, "crps"), c.fitc_eval(theta, U, "logs")))
        c.close()
    assert outs[0][0][0] == outs[1][0][0] and np.array_equal(outs[0][0][1], outs[1][0][1])
    assert outs[0][1][0] == outs[1][1][0] and np.array_equal(outs[0][1][2], outs[1][1][2])


def test_full_many_input_dimensions(ctx):
    """D = 40 inputs (the Gram and contraction kernels stage D x 128 panels in shared memory; limit 64)."""
    from oracle import gp_oracle as O
    rng = np.random.default_rng(77)
    n, d = 300, 40
    X = rng.standard_normal((n, d))
    y = rng.standard_normal((n, 1))
    theta = np.concatenate([[0.1], np.log(rng.uniform(3.0, 6.0, d)), [-1.5]])
    ctx.set_data(_dev(X), _dev(y))
    for score in ("crps", "logs", "nlml"):
        val, grad = ctx.full_eval(theta, score)
        oval, og = O.full_obj_grad(X, y, theta, O.SCORES[score])
        assert abs(val - oval) <= 1e-8 * abs(oval), score
        assert relerr(grad, og) <= 1e-6, score
