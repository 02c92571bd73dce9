"""Path-level parity of the FITC CUDA path (-m gpu) through the C-ABI against the
reference-generated goldens (dense big_Q path, K20:222-236 etc.) and the CPU oracle.
Tolerances: objective 1e-8 relative, gradients (theta and inducing inputs) 1e-6 relative."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, grad_vector, relerr

pytestmark = pytest.mark.gpu

OBJ_TOL = 1e-8
GRAD_TOL = 1e-6


@pytest.fixture(scope="module")
def ctx():
    from gpscore_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", golden_names(("c2", "c4")))
@pytest.mark.parametrize("score", ["crps", "logs", "nlml"])
def test_fitc_obj_grad_vs_reference_golden(ctx, name, score):
    g = load_golden(name)
    ctx.set_data(_dev(g["X"]), _dev(g["y"]))
    val, grad, gU = ctx.fitc_eval(g["theta"], g["U"], score)
    assert abs(val - g["obj_" + score]) <= OBJ_TOL * abs(g["obj_" + score])
    ref = grad_vector(g, score)
    if int(g["d_b"]) == 1 and g["X"].shape[1] > 1:
        grad = np.concatenate([[grad[0]], [grad[1:-1].sum()], [grad[-1]]])
    assert relerr(grad, ref) <= GRAD_TOL
    assert relerr(gU, g["grad_u_" + score]) <= GRAD_TOL
    if score != "nlml":
        mu, s2 = ctx.fitc_loo()
        assert relerr(mu.cpu().numpy(), g["loo_mean_" + score]) <= 1e-8
        assert relerr(s2.cpu().numpy(), g["loo_var_" + score]) <= 1e-8


@pytest.mark.parametrize("name", golden_names(("c2", "c4")))
def test_fitc_predict_and_metrics_vs_golden(ctx, name):
    g = load_golden(name)
    ctx.set_data(_dev(g["X"]), _dev(g["y"]))
    mean, var = ctx.fitc_predict(g["theta"], g["U"], _dev(g["Xs"]))
    assert relerr(mean.cpu().numpy(), g["pred_mean"]) <= 1e-8
    assert relerr(var.cpu().numpy(), g["pred_var"]) <= 1e-7
    m = ctx.test_metrics(mean, var, _dev(g["ys"]), _dev(g["y"]))
    for k in ("mse", "smse", "logs", "crps", "msll", "coverage"):
        assert abs(m[k] - g["m_" + k]) <= 1e-7 * max(1.0, abs(g["m_" + k])), k


@pytest.mark.parametrize("m_ind", [1, 8, 9, 20, 32])
def test_fitc_all_m_paddings_vs_oracle(ctx, m_ind):
    from gpscore_b200 import synth
    from oracle import gp_oracle as O
    X, y = synth.kin40k_like(700, seed=40)
    theta = synth.hyper_point("P1", seed=41)
    U = synth.inducing_init(m_ind, seed=42) * 2 - 1
    ctx.set_data(_dev(X), _dev(y))
    for score in ("crps", "logs", "nlml"):
        val, grad, gU = ctx.fitc_eval(theta, U, score)
        oval, og, ogU, _ = O.fitc_obj_grad(X, y, U, theta, O.SCORES[score])
        assert abs(val - oval) <= OBJ_TOL * abs(oval), score
        assert relerr(grad, og) <= GRAD_TOL, score
        assert relerr(gU, ogU) <= GRAD_TOL, score


def test_fitc_row_sharded_equals_single(ctx):
    """Two contexts each holding half of the rows (the multi-GPU layout, emulated on one GPU with a
    python sum standing in for the NCCL all-reduce) reproduce the single-context result."""
    from gpscore_b200 import api, synth
    X, y = synth.kin40k_like(1001, seed=50)
    theta = synth.hyper_point("P2")
    U = synth.inducing_init(20, seed=51)
    ctx.set_data(_dev(X), _dev(y))
    parts = []
    cuts = [0, 400, 1001]
    for r in range(2):
        c = api.Context(0)
        c.set_data(_dev(X[cuts[r]:cuts[r + 1]]), _dev(y[cuts[r]:cuts[r + 1]]))
        parts.append(c)
    for score in ("crps", "nlml"):
        ref = ctx.fitc_eval(theta, U, score)
        # lock-step emulation of the three all-reduces
        import ctypes as C
        from gpscore_b200.api import _dp, _host_vec
        from gpscore_b200 import lib as L
        th, Uh = np.ascontiguousarray(theta), _host_vec(U)
        l1, l2, l3 = ctx.fitc_acc_len(20)
        accs = [[torch.zeros(n, dtype=torch.float64, device="cuda") for n in (l1, l2, l3)] for _ in parts]
        for c in parts:
            c._check(c._lib.gps_fitc_begin(c._h, _dp(th), _dp(Uh), 20, 1e-3, L.SCORES[score], 1001))
        for k, fn in enumerate(("gps_fitc_pass1", "gps_fitc_pass2", "gps_fitc_pass3")):
            for c, a in zip(parts, accs):
                if k == 0:
                    c._check(getattr(c._lib, fn)(c._h, a[0].data_ptr()))
                else:
                    c._check(getattr(c._lib, fn)(c._h, a[k - 1].data_ptr(), a[k].data_ptr()))
            tot = accs[0][k] + accs[1][k]
            for a in accs:
                a[k].copy_(tot)
        outs = []
        for c, a in zip(parts, accs):
            obj, g, gU = np.zeros(1), np.zeros(10), np.zeros(160)
            c._check(c._lib.gps_fitc_finish(c._h, a[1].data_ptr(), a[2].data_ptr(), _dp(obj), _dp(g), _dp(gU)))
            outs.append((obj[0], g, gU.reshape(20, 8)))
        for o in outs:
            assert abs(o[0] - ref[0]) <= 1e-12 * abs(ref[0])
            assert relerr(o[1], ref[1]) <= 1e-10
            assert relerr(o[2], ref[2]) <= 1e-10
    for c in parts:
        c.close()


def test_fitc_full_size_finite_difference(ctx):
    """N = 10 000, M = 20 (BASELINE's FITC size): directional finite difference of the objective
    against the analytic gradient over all 170 parameters."""
    from gpscore_b200 import synth
    X, y = synth.kin40k_like(10000, seed=0)
    theta = synth.hyper_point("P1")
    U = synth.inducing_init(20)
    ctx.set_data(_dev(X), _dev(y))
    rng = np.random.default_rng(6)
    vt = rng.standard_normal(theta.size)
    vU = rng.standard_normal(U.shape)
    nrm = np.sqrt((vt ** 2).sum() + (vU ** 2).sum())
    vt, vU = vt / nrm, vU / nrm
    for score in ("crps", "logs", "nlml"):
        val, g, gU = ctx.fitc_eval(theta, U, score)
        h = 1e-5
        fp = ctx.fitc_eval(theta + h * vt, U + h * vU, score)[0]
        fm = ctx.fitc_eval(theta - h * vt, U - h * vU, score)[0]
        fd = (fp - fm) / (2 * h)
        an = g @ vt + (gU * vU).sum()
        assert abs(fd - an) <= 1e-6 * max(1.0, abs(fd)), (score, fd, an)


def test_fitc_script_loop(ctx):
    """K20:219-251 with the objective statements replaced by the fused call; 5 GD steps including the
    inducing-input update equal the oracle-driven loop."""
    from gpscore_b200 import api, synth
    from oracle import gp_oracle as O
    X, y = synth.kin40k_like(300, seed=5)
    train_x, train_y = _dev(X), _dev(y)
    theta = synth.hyper_point("P1")
    U0 = synth.inducing_init(20)
    para_l = torch.tensor(theta[1:-1].reshape(1, 8), requires_grad=True)
    para_k = torch.tensor([theta[0]], requires_grad=True)
    para_noise = torch.tensor([theta[-1]], requires_grad=True)
    inducing_x = torch.tensor(U0, requires_grad=True)
    th, U = theta.copy(), U0.copy()
    for i in range(5):
        learning_rate = 1
        CRPS_ave = api.fitc_loo_objective(train_x, train_y, inducing_x, para_k, para_l, para_noise, "crps", ctx=ctx)
        CRPS_ave.backward()
        with torch.no_grad():
            para_l -= learning_rate * para_l.grad
            para_k -= learning_rate * para_k.grad
            para_noise -= learning_rate * para_noise.grad
            inducing_x -= learning_rate * inducing_x.grad
            para_l.grad.zero_()
            para_k.grad.zero_()
            para_noise.grad.zero_()
            inducing_x.grad.zero_()
        _, g, gU, _ = O.fitc_obj_grad(X, y, U, th, O.SCORE_CRPS)
        th = th - learning_rate * g
        U = U - learning_rate * gU
    got = np.concatenate([para_k.detach().numpy(), para_l.detach().numpy().ravel(), para_noise.detach().numpy()])
    assert relerr(got, th) <= 1e-6
    assert relerr(inducing_x.detach().numpy(), U) <= 1e-6


@pytest.mark.parametrize("m_ind", [5, 16, 20, 24, 31, 32])
def test_fitc_row_pass_formulations_agree(ctx, m_ind):
    """thread-per-row passes (variant 0), the tile/DMMA formulation (variant 1) and the fused three-kernel
    path (variant 2, default; M = 32 falls back to the tile kernels) are the same computation: objective and
    all gradients agree to rounding."""
    from gpscore_b200 import synth
    X, y = synth.kin40k_like(900, seed=60)
    theta = synth.hyper_point("P2")
    U = synth.inducing_init(m_ind, seed=61)
    ctx.set_data(_dev(X), _dev(y))
    try:
        res = {}
        for variant in (0, 1, 2):
            ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 2, variant))
            res[variant] = [ctx.fitc_eval(theta, U, s) for s in ("crps", "logs", "nlml")]
        for other in (1, 2):
            for a, b in zip(res[0], res[other]):
                assert abs(a[0] - b[0]) <= 1e-11 * abs(a[0]), other
                assert relerr(a[1], b[1]) <= 1e-9 and relerr(a[2], b[2]) <= 1e-9, other
    finally:
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 2, 2))


@pytest.mark.parametrize("kind", ["dss", "kc"])
@pytest.mark.parametrize("name", [n for n in golden_names(("c4",)) if "ragged" not in n])
def test_fitc_dss_vs_reference_golden(ctx, name, kind):
    """FITC 4-fold DSS (K20:538-587) and block-CRPS kc (K20:669-720) objectives + gradients in Woodbury
    form vs the reference's autograd."""
    g = load_golden(name)
    ctx.set_data(_dev(g["X"]), _dev(g["y"]))
    val, grad, gU = ctx.fitc_eval(g["theta"], g["U"], kind)
    assert abs(val - g["obj_" + kind]) <= OBJ_TOL * abs(g["obj_" + kind])
    ref = grad_vector(g, kind)
    if int(g["d_b"]) == 1 and g["X"].shape[1] > 1:
        grad = np.concatenate([[grad[0]], [grad[1:-1].sum()], [grad[-1]]])
    assert relerr(grad, ref) <= GRAD_TOL
    assert relerr(gU, g["grad_u_" + kind]) <= GRAD_TOL


@pytest.mark.parametrize("kind", ["dss", "kc"])
@pytest.mark.parametrize("m_ind,n", [(3, 128), (20, 2000), (32, 1204)])
def test_fitc_dss_vs_oracle(ctx, m_ind, n, kind):
    from gpscore_b200 import synth, lib as L
    from oracle import woodbury as WB
    X, y = synth.kin40k_like(n, seed=70)
    theta = synth.hyper_point("P2")
    U = synth.inducing_init(m_ind, seed=71) * 2 - 1
    ctx.set_data(_dev(X), _dev(y))
    val, grad, gU = ctx.fitc_eval(theta, U, kind)
    oval, og, ogU = WB.fitc_block_obj_grad(X, y, U, theta, kind)
    assert abs(val - oval) <= OBJ_TOL * abs(oval)
    assert relerr(grad, og) <= GRAD_TOL and relerr(gU, ogU) <= GRAD_TOL
    ctx.set_data(_dev(X[:n - 1]), _dev(y[:n - 1]))      # 4 does not divide N: refused like the script's fold code
    with pytest.raises(L.GpsError):
        ctx.fitc_eval(theta, U, kind)
