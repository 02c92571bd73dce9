"""Path-level parity of the full-GP CUDA path (-m gpu), through the C-ABI, against
 (a) the reference-generated goldens (tests/golden/*.npz) and (b) the CPU oracle on seeded inputs.
Tolerances are BASELINE.json's: objective 1e-8 relative, gradient 1e-6 relative."""
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


@pytest.mark.parametrize("name", golden_names(("c1", "c3")))
@pytest.mark.parametrize("score", ["crps", "logs", "nlml"])
def test_full_obj_grad_vs_reference_golden(ctx, name, score):
    g = load_golden(name)
    ctx.set_data(_dev(g["X"]), _dev(g["y"]))
    val, grad = ctx.full_eval(g["theta"], score)
    assert abs(val - g["obj_" + score]) <= OBJ_TOL * abs(g["obj_" + score])
    ref = grad_vector(g, score)
    if int(g["d_b"]) == 1 and g["X"].shape[1] > 1:
        grad = np.concatenate([[grad[0]], [grad[1:-1].sum()], [grad[-1]]])
    assert relerr(grad, ref) <= GRAD_TOL
    # objective-only call gives the same value
    val2, none = ctx.full_eval(g["theta"], score, grad=False)
    assert none is None and abs(val2 - val) <= 1e-12 * abs(val)
    if score != "nlml":
        mu, s2 = ctx.full_loo()
        assert relerr(mu.cpu().numpy(), g["loo_mean_" + score]) <= 1e-8
        assert relerr(s2.cpu().numpy(), g["loo_var_" + score]) <= 1e-8


@pytest.mark.parametrize("name", golden_names(("c1", "c3")))
def test_full_predict_and_metrics_vs_golden(ctx, name):
    g = load_golden(name)
    ctx.set_data(_dev(g["X"]), _dev(g["y"]))
    mean, var = ctx.full_predict(g["theta"], _dev(g["Xs"]))
    assert relerr(mean.cpu().numpy(), g["pred_mean"]) <= 1e-8
    assert relerr(var.cpu().numpy(), g["pred_var"]) <= 1e-8
    m = ctx.test_metrics(mean, var, _dev(g["ys"]), _dev(g["y"]))
    for k in ("mse", "smse", "logs", "crps", "msll", "coverage"):
        assert abs(m[k] - g["m_" + k]) <= 1e-8 * max(1.0, abs(g["m_" + k])), k


@pytest.mark.parametrize("n", [2000])
@pytest.mark.parametrize("point", ["P1", "P2"])
def test_full_obj_grad_vs_oracle_scaled(ctx, n, point):
    from gpscore_b200 import synth
    from oracle import gp_oracle as O
    X, y = synth.kin40k_like(n, seed=30)
    theta = synth.hyper_point(point)
    ctx.set_data(_dev(X), _dev(y))
    for score in ("crps", "logs", "nlml"):
        val, grad = ctx.full_eval(theta, score)
        oval, ograd = O.full_obj_grad(X, y, theta, O.SCORES[score])
        assert abs(val - oval) <= OBJ_TOL * abs(oval), score
        assert relerr(grad, ograd) <= GRAD_TOL, score


def test_full_size_independent_properties(ctx):
    """At BASELINE's full size (N = 10 000) the oracle takes minutes, so the check is through
    properties: (1) the gradient agrees with a central finite difference of the objective along a
    random direction, (2) LOO quantities satisfy K^-1 identities on a probe vector."""
    from gpscore_b200 import synth
    n = 10000
    X, y = synth.kin40k_like(n, seed=0)
    theta = synth.hyper_point("P1")
    ctx.set_data(_dev(X), _dev(y))
    rng = np.random.default_rng(4)
    v = rng.standard_normal(theta.size)
    v /= np.linalg.norm(v)
    for score in ("crps", "nlml"):
        val, grad = ctx.full_eval(theta, score)
        h = 1e-5
        fp, _ = ctx.full_eval(theta + h * v, score, grad=False)
        fm, _ = ctx.full_eval(theta - h * v, score, grad=False)
        fd = (fp - fm) / (2 * h)
        assert abs(fd - grad @ v) <= 1e-6 * max(1.0, abs(fd)), (score, fd, grad @ v)
    # LOO variance of the noisy target is below the prior variance and positive
    ctx.full_eval(theta, "crps")
    mu, s2 = ctx.full_loo()
    assert float(s2.min()) > 0 and float(s2.max()) <= np.exp(theta[0]) + np.exp(theta[-1]) + 1e-9


def test_script_loop_with_fused_objective(ctx):
    """The optimiser loop of KF:237-260 with the six objective statements replaced by the fused
    call: parameters after 5 steps equal those of the oracle-driven loop."""
    from gpscore_b200 import api, synth
    from oracle import gp_oracle as O
    X, y = synth.kin40k_like(300, seed=5)
    train_x, train_y = _dev(X), _dev(y)
    torch.manual_seed(0)
    para_l = torch.rand(1, 8, dtype=torch.float64, requires_grad=True)
    para_k = torch.tensor([1.0], dtype=torch.float64, requires_grad=True)
    para_noise = torch.tensor([1.0], dtype=torch.float64, requires_grad=True)
    theta = np.concatenate([[1.0], para_l.detach().numpy().ravel(), [1.0]])
    for i in range(5):
        learning_rate = 1
        CRPS_ave = api.full_loo_objective(train_x, train_y, para_k, para_l, para_noise, "crps", ctx=ctx)
        CRPS_ave.backward()
        with torch.no_grad():
            para_l -= learning_rate * para_l.grad
            para_k -= learning_rate * para_k.grad
            para_noise -= learning_rate * para_noise.grad
            para_l.grad.zero_()
            para_k.grad.zero_()
            para_noise.grad.zero_()
        _, g = O.full_obj_grad(X, y, theta, O.SCORE_CRPS)
        theta = theta - learning_rate * g
    got = np.concatenate([para_k.detach().numpy(), para_l.detach().numpy().ravel(), para_noise.detach().numpy()])
    assert relerr(got, theta) <= 1e-6


def test_reference_named_helpers(ctx):
    from gpscore_b200 import api
    from oracle import gp_oracle as O
    rng = np.random.default_rng(9)
    n, t = 150, 40
    X, Xs = rng.standard_normal((n, 3)), rng.standard_normal((t, 3))
    y = rng.standard_normal((n, 1))
    a, b, c = 0.2, np.array([[0.1, 0.3, -0.2]]), -1.5
    K = O.ARD(X, X, a, b)
    k1 = O.ARD(Xs, X, a, b)
    k3 = O.ARD(Xs, Xs, a, b)
    sol = api.chol_solve(_dev(y), _dev(K + np.exp(c) * np.eye(n)))
    assert relerr(sol.cpu().numpy(), O.chol_solve(y, K + np.exp(c) * np.eye(n))) <= 1e-9
    mean, cov = api.cal_mean_and_cov(_dev(k1), _dev(K), _dev(k3), t, n, _dev(y), sigma_noise_sq=np.exp(c))
    rm, rc = O.cal_mean_and_cov(k1, K, k3, t, n, y, np.exp(c))
    assert relerr(mean.cpu().numpy(), rm) <= 1e-9 and relerr(cov.cpu().numpy(), rc) <= 1e-9
    m, v, yy = rng.standard_normal((t, 1)), rng.random((t, 1)) + 0.1, rng.standard_normal((t, 1))
    assert abs(float(api.crps(_dev(m), _dev(v), _dev(yy))) - O.crps(m, v, yy)) <= 1e-13
    assert abs(float(api.logs(_dev(m), _dev(v), _dev(yy))) - O.logs(m, v, yy)) <= 1e-13
    assert abs(float(api.SMSE(_dev(m), _dev(yy), _dev(y))) - O.SMSE(m, yy, y)) <= 1e-13
    assert abs(float(api.trivial_loss(_dev(m), _dev(v), _dev(yy), _dev(y))) - O.trivial_loss(m, v, yy, y)) <= 1e-12


def test_descend_equals_python_loop(ctx):
    """gps_full_descend / gps_fitc_descend (the scripts' optimiser loop in one C-ABI call) give the
    same trajectory as stepping from Python."""
    from gpscore_b200 import synth
    X, y = synth.kin40k_like(400, seed=8)
    ctx.set_data(_dev(X), _dev(y))
    theta = synth.hyper_point("P1")
    th, trace = ctx.full_descend(theta, "crps", 1.0, 6)
    ref = theta.copy()
    for i in range(6):
        v, g = ctx.full_eval(ref, "crps")
        assert abs(v - trace[i]) <= 1e-12 * abs(v)
        ref = ref - 1.0 * g
    assert relerr(th, ref) <= 1e-12
    U = synth.inducing_init(20)
    th, Uo, trace = ctx.fitc_descend(theta, U, "logs", 0.2, 0.1, 5)
    ref, Ur = theta.copy(), U.copy()
    for i in range(5):
        v, g, gU = ctx.fitc_eval(ref, Ur, "logs")
        assert abs(v - trace[i]) <= 1e-12 * abs(v)
        ref, Ur = ref - 0.2 * g, Ur - 0.1 * gU
    assert relerr(th, ref) <= 1e-12 and relerr(Uo, Ur) <= 1e-12
    # the same loop beyond 32 inducing points (matrix form of the FITC passes)
    U = X[:48] + 0.01
    th, Uo, trace = ctx.fitc_descend(theta, U, "crps", 0.2, 0.1, 3)
    ref, Ur = theta.copy(), U.copy()
    for i in range(3):
        v, g, gU = ctx.fitc_eval(ref, Ur, "crps")
        assert abs(v - trace[i]) <= 1e-12 * abs(v)
        ref, Ur = ref - 0.2 * g, Ur - 0.1 * gU
    assert relerr(th, ref) <= 1e-12 and relerr(Uo, Ur) <= 1e-12


@pytest.mark.parametrize("name", [n for n in golden_names(("c3",)) if "ragged" not in n])
def test_full_dss_vs_reference_golden(ctx, name):
    """4-fold block-LOO DSS objective + gradient (KF:499-543) against the reference's autograd."""
    g = load_golden(name)
    ctx.set_data(_dev(g["X"]), _dev(g["y"]))
    val, grad = ctx.full_eval(g["theta"], "dss")
    assert abs(val - g["obj_dss"]) <= OBJ_TOL * abs(g["obj_dss"])
    assert relerr(grad, grad_vector(g, "dss")) <= GRAD_TOL
    val2, _ = ctx.full_eval(g["theta"], "dss", grad=False)
    assert abs(val2 - val) <= 1e-12 * abs(val)


def test_full_dss_vs_oracle_and_fold_rule(ctx):
    from gpscore_b200 import synth, lib as L
    from oracle import gp_oracle as O
    X, y = synth.kin40k_like(2000, seed=31)          # folds of 500: not tile aligned
    theta = synth.hyper_point("P2")
    ctx.set_data(_dev(X), _dev(y))
    val, grad = ctx.full_eval(theta, "dss")
    oval, ograd = O.full_dss_obj_grad(X, y, theta)
    assert abs(val - oval) <= OBJ_TOL * abs(oval)
    assert relerr(grad, ograd) <= GRAD_TOL
    # the reference's fold code only works for 4 | N (KF:521-530): refused, not silently wrong
    ctx.set_data(_dev(X[:1999]), _dev(y[:1999]))
    with pytest.raises(L.GpsError):
        ctx.full_eval(theta, "dss")


def test_factorisation_schedules_are_bit_identical(ctx):
    """The shipped schedule (POTRF lanes + TRTRI merges overlapped on priority streams), POTRF-then-TRTRI and the
    fully serialised timing mode run the same task lists: objective and gradient agree bit for bit, and
    repeated evaluations of the overlapped schedule are reproducible (no race between the lanes)."""
    from gpscore_b200 import synth
    X, y = synth.kin40k_like(3000, seed=12)
    theta = synth.hyper_point("P2")
    ctx.set_data(_dev(X), _dev(y))
    try:
        res = {}
        for mode in (1, 0, 2, 1):
            ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 4, mode))
            for score in ("crps", "nlml"):
                for rep in range(3):
                    v, g = ctx.full_eval(theta, score)
                    key = res.setdefault(score, (v, g))
                    assert v == key[0] and np.array_equal(g, key[1]), (mode, score, rep)
    finally:
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 4, 1))


def test_gemm_policies_do_not_change_results(ctx):
    """The row-strip policies (chain launches, knob 7; automatic for under-filled launches, knob 8) split a tile
    over more CTAs but keep the k-order per output element: objective and gradient are bit-identical."""
    from gpscore_b200 import synth
    X, y = synth.kin40k_like(1700, seed=13)
    theta = synth.hyper_point("P1")
    ctx.set_data(_dev(X), _dev(y))
    try:
        ref = None
        for chain, auto in ((16, 1), (32, 1), (0, 1), (16, 0), (0, 0)):
            ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 7, chain))
            ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 8, auto))
            v, g = ctx.full_eval(theta, "crps")
            ref = ref or (v, g)
            assert v == ref[0] and np.array_equal(g, ref[1]), (chain, auto)
    finally:
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 7, 16))
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 8, 1))


def test_lane_trace_reports_every_outer_step(ctx):
    """gps_dbg_trace: one entry per lane and outer step, times non-decreasing within a lane."""
    import ctypes as C
    from gpscore_b200 import synth
    X, y = synth.kin40k_like(5000, seed=14)          # 40 tiles: at least three outer steps for any block width <= 13
    ctx.set_data(_dev(X), _dev(y))
    try:
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 6, 1))
        ctx.full_eval(synth.hyper_point("P1"), "crps")
        codes, ms = (C.c_int * 512)(), (C.c_double * 512)()
        n = ctx._lib.gps_dbg_trace(ctx._h, 512, codes, ms)
        assert 0 < n < 512
        lanes = {}
        for i in range(n):
            lanes.setdefault(codes[i] // 1000, []).append((codes[i] % 1000, ms[i]))
        steps = [s for s, _ in lanes[1]]
        assert len(steps) >= 3 and steps == list(range(len(steps)))
        for lane in (2, 3, 4):
            assert [s for s, _ in lanes[lane]] == steps, lane
        for lane in (1, 2, 3, 4):
            times = [t for _, t in lanes[lane]]
            assert all(b >= a for a, b in zip(times, times[1:]))
        assert [k for k, _ in lanes[5]] == list(range(40))      # one mark per 128-wide step of the chain
    finally:
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 6, 0))
