"""FITC with M > 32 inducing points (-m gpu): the matrix form of csrc/gps_fitc_large.cu through the
C-ABI against the CPU oracle and — with the switch-over lowered so that M = 20 runs it too — against
the reference-generated goldens.  Tolerances: objective 1e-8 relative, gradients 1e-6 relative."""
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


@pytest.fixture()
def forced(ctx):
    """Lower the switch-over so every M runs the matrix form."""
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 3, 1))
    yield ctx
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 3, 33))


@pytest.mark.parametrize("name", golden_names(("c2", "c4")))
@pytest.mark.parametrize("score", ["crps", "logs", "nlml"])
def test_matrix_form_vs_reference_golden(forced, name, score):
    g = load_golden(name)
    forced.set_data(_dev(g["X"]), _dev(g["y"]))
    val, grad, gU = forced.fitc_eval(g["theta"], g["U"], score)
    assert abs(val - g["obj_" + score]) <= OBJ_TOL * abs(g["obj_" + score])
    ref = grad_vector(g, score)
    if int(g["d_b"]) == 1 and g["X"].shape[1] > 1:
        grad = np.concatenate([[grad[0]], [grad[1:-1].sum()], [grad[-1]]])
    assert relerr(grad, ref) <= GRAD_TOL
    assert relerr(gU, g["grad_u_" + score]) <= GRAD_TOL
    if score != "nlml":
        mu, s2 = forced.fitc_loo()
        assert relerr(mu.cpu().numpy(), g["loo_mean_" + score]) <= 1e-8
        assert relerr(s2.cpu().numpy(), g["loo_var_" + score]) <= 1e-8


@pytest.mark.parametrize("name", golden_names(("c4",))[:2])
def test_matrix_form_predict_vs_golden(forced, name):
    g = load_golden(name)
    forced.set_data(_dev(g["X"]), _dev(g["y"]))
    mean, var = forced.fitc_predict(g["theta"], g["U"], _dev(g["Xs"]), force_matrix_form=True)
    assert relerr(mean.cpu().numpy(), g["pred_mean"]) <= 1e-8
    assert relerr(var.cpu().numpy(), g["pred_var"]) <= 1e-7


@pytest.mark.parametrize("m_ind,n,d", [(33, 700, 8), (128, 1000, 8), (200, 1500, 8), (300, 900, 3), (64, 777, 12)])
def test_large_m_vs_oracle(ctx, m_ind, n, d):
    from oracle import woodbury as Wd
    rng = np.random.default_rng(100 + m_ind)
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(X @ rng.standard_normal(d)) + 0.1 * rng.standard_normal(n)
    U = rng.uniform(-1, 1, (m_ind, d))
    theta = np.concatenate([[0.3], np.log(rng.uniform(0.8, 2.0, d)), [-2.0]])
    ctx.set_data(_dev(X), _dev(y))
    for score in ("crps", "logs", "nlml"):
        from oracle import gp_oracle as O
        val, grad, gU = ctx.fitc_eval(theta, U, score)
        oval, og, ogU, lm, lv = Wd.fitc_obj_grad(X, y, U, theta, O.SCORES[score])
        assert abs(val - oval) <= OBJ_TOL * abs(oval), score
        assert relerr(grad, og) <= GRAD_TOL, score
        assert relerr(gU, ogU) <= GRAD_TOL, score
        mu, s2 = ctx.fitc_loo()
        assert relerr(mu.cpu().numpy().ravel(), lm) <= 1e-7
        assert relerr(s2.cpu().numpy().ravel(), lv) <= 1e-7
        val2, _, _ = ctx.fitc_eval(theta, U, score)
        assert val2 == val                          # fixed reduction order: bit-reproducible


def test_large_m_predict_vs_oracle(ctx):
    from oracle import woodbury as Wd
    rng = np.random.default_rng(7)
    n, d, m_ind, t = 1100, 8, 96, 333
    X = rng.uniform(-1, 1, (n, d))
    y = np.cos(X @ rng.standard_normal(d)) + 0.05 * rng.standard_normal(n)
    Xs = rng.uniform(-1, 1, (t, d))
    U = rng.uniform(-1, 1, (m_ind, d))
    theta = np.concatenate([[0.1], np.log(rng.uniform(0.8, 2.0, d)), [-3.0]])
    ctx.set_data(_dev(X), _dev(y))
    mean, var = ctx.fitc_predict(theta, U, _dev(Xs))
    om, ov = Wd.fitc_predict(X, y, U, Xs, theta)
    assert relerr(mean.cpu().numpy(), om) <= 1e-7
    assert relerr(var.cpu().numpy(), ov) <= 1e-7


def test_large_m_objective_only_and_errors(ctx):
    from gpscore_b200 import lib as L
    rng = np.random.default_rng(9)
    X = rng.uniform(-1, 1, (600, 8))
    y = rng.standard_normal(600)
    U = rng.uniform(-1, 1, (40, 8))
    theta = np.zeros(10)
    ctx.set_data(_dev(X), _dev(y))
    full = ctx.fitc_eval(theta, U, "crps")
    obj = np.zeros(1)
    from gpscore_b200.api import _dp, _host_vec
    ctx._check(ctx._lib.gps_fitc_eval(ctx._h, _dp(theta), _dp(_host_vec(U)), 40, 1e-3, 0, _dp(obj), None, None))
    assert obj[0] == full[0]
    with pytest.raises(L.GpsError):                 # the fold code needs 4 | N (K20:541-543)
        ctx.set_data(_dev(X[:598]), _dev(y[:598]))
        ctx.fitc_eval(theta, U, "dss")


@pytest.mark.parametrize("score", ["crps", "logs", "nlml"])
def test_large_m_row_sharded_equals_single(ctx, score):
    """Three contexts holding uneven row blocks (the multi-GPU layout, emulated on one GPU with a python
    sum standing in for the NCCL all-reduce) reproduce the single-context matrix-form result."""
    from gpscore_b200 import api, lib as L
    from gpscore_b200.api import _dp, _host_vec
    rng = np.random.default_rng(21)
    n, d, m_ind = 1501, 8, 96
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(X @ rng.standard_normal(d)) + 0.1 * rng.standard_normal(n)
    U = rng.uniform(-1, 1, (m_ind, d))
    theta = np.concatenate([[0.2], np.log(rng.uniform(0.8, 2.0, d)), [-2.5]])
    ctx.set_data(_dev(X), _dev(y))
    ref = ctx.fitc_eval(theta, U, score)
    cuts = [0, 300, 1100, n]
    parts = []
    for r in range(3):
        c = api.Context(0)
        c.set_data(_dev(X[cuts[r]:cuts[r + 1]]), _dev(y[cuts[r]:cuts[r + 1]]))
        parts.append(c)
    th, Uh = np.ascontiguousarray(theta), _host_vec(U)
    lens = ctx.fitc_acc_len(m_ind)
    accs = [[torch.zeros(k, dtype=torch.float64, device="cuda") for k in lens] for _ in parts]
    for c in parts:
        c._check(c._lib.gps_fitc_begin(c._h, _dp(th), _dp(Uh), m_ind, 1e-3, L.SCORES[score], n))
    for k, fn in enumerate(("gps_fitc_pass1", "gps_fitc_pass2", "gps_fitc_pass3")):
        for c, a in zip(parts, accs):
            if k == 0:
                c._check(getattr(c._lib, fn)(c._h, a[0].data_ptr()))
            else:
                c._check(getattr(c._lib, fn)(c._h, a[k - 1].data_ptr(), a[k].data_ptr()))
        tot = accs[0][k] + accs[1][k] + accs[2][k]
        for a in accs:
            a[k].copy_(tot)
    for c, a in zip(parts, accs):
        obj, g, gU = np.zeros(1), np.zeros(d + 2), np.zeros(m_ind * d)
        c._check(c._lib.gps_fitc_finish(c._h, a[1].data_ptr(), a[2].data_ptr(), _dp(obj), _dp(g), _dp(gU)))
        assert abs(obj[0] - ref[0]) <= 1e-11 * abs(ref[0])
        assert relerr(g, ref[1]) <= 1e-9
        assert relerr(gU.reshape(m_ind, d), ref[2]) <= 1e-9
    # the public sharded entry point with a one-rank "all-reduce"
    one = ctx.fitc_eval_sharded(theta, U, score, n, lambda t: t)
    assert abs(one[0] - ref[0]) <= 1e-13 * abs(ref[0]) and relerr(one[1], ref[1]) <= 1e-12
    for c in parts:
        c.close()


def test_large_m_finite_difference(ctx):
    """M = 256, N = 20 000: directional finite difference over theta and U."""
    from gpscore_b200 import synth
    X, y = synth.kin40k_like(20000, seed=3)
    theta = synth.hyper_point("P1")
    rng = np.random.default_rng(11)
    U = X[rng.choice(20000, 256, replace=False)] + 0.01 * rng.standard_normal((256, 8))
    ctx.set_data(_dev(X), _dev(y))
    vt = rng.standard_normal(theta.size)
    vu = rng.standard_normal(U.shape)
    for score in ("crps", "nlml"):
        _, g, gU = ctx.fitc_eval(theta, U, score)
        an = g @ vt + np.sum(gU * vu)
        h = 1e-5
        fp = ctx.fitc_eval(theta + h * vt, U + h * vu, score)[0]
        fm = ctx.fitc_eval(theta - h * vt, U - h * vu, score)[0]
        fd = (fp - fm) / (2 * h)
        assert abs(fd - an) <= 2e-5 * max(abs(an), 1e-3), (score, fd, an)


@pytest.mark.parametrize("n,m_ind,d", [(30, 40, 2), (129, 33, 1), (128, 128, 5), (40, 129, 16)])
def test_large_m_tiny_and_boundary_shapes(ctx, n, m_ind, d):
    """Fewer rows than inducing points, sizes on / next to the 128 padding boundary, D = 1 and D = 16."""
    from oracle import gp_oracle as O
    from oracle import woodbury as Wd
    rng = np.random.default_rng(1000 + n + m_ind)
    X = rng.uniform(-1, 1, (n, d))
    y = rng.standard_normal(n)
    U = rng.uniform(-1, 1, (m_ind, d))
    theta = np.concatenate([[0.2], np.log(rng.uniform(0.7, 1.5, d)), [-1.0]])
    ctx.set_data(_dev(X), _dev(y))
    for score in ("crps", "logs", "nlml"):
        val, grad, gU = ctx.fitc_eval(theta, U, score)
        oval, og, ogU, _, _ = Wd.fitc_obj_grad(X, y, U, theta, O.SCORES[score])
        assert abs(val - oval) <= OBJ_TOL * abs(oval), score
        assert relerr(grad, og) <= GRAD_TOL, score
        assert relerr(gU, ogU) <= GRAD_TOL, score


# ---- block objectives in the matrix form (4-fold DSS K20:538-587, block CRPS "kc" K20:669-720) ---------------

@pytest.mark.parametrize("kind", ["dss", "kc"])
@pytest.mark.parametrize("name", [n for n in golden_names(("c4",)) if "ragged" not in n])
def test_matrix_form_block_objectives_vs_reference_golden(forced, name, kind):
    """Switch-over lowered: the goldens' M = 20 runs the matrix form; values and gradients are the reference's
    own autograd results."""
    g = load_golden(name)
    forced.set_data(_dev(g["X"]), _dev(g["y"]))
    val, grad, gU = forced.fitc_eval(g["theta"], g["U"], kind)
    assert abs(val - g["obj_" + kind]) <= OBJ_TOL * abs(g["obj_" + kind])
    ref = grad_vector(g, kind)
    if int(g["d_b"]) == 1 and g["X"].shape[1] > 1:
        grad = np.concatenate([[grad[0]], [grad[1:-1].sum()], [grad[-1]]])
    assert relerr(grad, ref) <= GRAD_TOL
    assert relerr(gU, g["grad_u_" + kind]) <= GRAD_TOL


@pytest.mark.parametrize("kind", ["dss", "kc"])
@pytest.mark.parametrize("m_ind,n,d", [(33, 700, 8), (64, 1204, 8), (130, 2000, 8), (200, 1500, 3), (40, 36, 2)])
def test_large_m_block_objectives_vs_oracle(ctx, m_ind, n, d, kind):
    """Fold boundaries off the 16-column k-granularity (n/4 = 175, 301, 375, 9) and off the 128-column tiles."""
    from oracle import woodbury as Wd
    rng = np.random.default_rng(300 + m_ind)
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(X @ rng.standard_normal(d)) + 0.1 * rng.standard_normal(n)
    U = rng.uniform(-1, 1, (m_ind, d))
    theta = np.concatenate([[0.3], np.log(rng.uniform(0.8, 2.0, d)), [-2.0]])
    ctx.set_data(_dev(X), _dev(y))
    val, grad, gU = ctx.fitc_eval(theta, U, kind)
    oval, og, ogU = Wd.fitc_block_obj_grad(X, y, U, theta, kind)
    assert abs(val - oval) <= OBJ_TOL * abs(oval)
    assert relerr(grad, og) <= GRAD_TOL
    assert relerr(gU, ogU) <= GRAD_TOL
    val2, grad2, gU2 = ctx.fitc_eval(theta, U, kind)
    assert val2 == val and np.array_equal(grad, grad2) and np.array_equal(gU, gU2)   # fixed reduction order
    # objective only (no gradient pointers) takes the short path and returns the same value
    from gpscore_b200.api import _dp, _host_vec
    from gpscore_b200 import lib as L
    obj = np.zeros(1)
    ctx._check(ctx._lib.gps_fitc_eval(ctx._h, _dp(theta), _dp(_host_vec(U)), m_ind, 1e-3, L.SCORES[kind], _dp(obj), None, None))
    assert abs(obj[0] - val) <= 1e-13 * abs(val)


@pytest.mark.parametrize("kind", ["dss", "kc"])
def test_matrix_form_block_objectives_equal_fused_small_m(ctx, kind):
    """The same objective through the M <= 32 row kernels and (switch-over lowered) the matrix form."""
    rng = np.random.default_rng(41)
    X = rng.uniform(-1, 1, (1600, 8))
    y = np.sin(X @ rng.standard_normal(8)) + 0.1 * rng.standard_normal(1600)
    U = rng.uniform(-1, 1, (24, 8))
    theta = np.concatenate([[0.2], np.log(rng.uniform(0.8, 2.0, 8)), [-2.5]])
    ctx.set_data(_dev(X), _dev(y))
    a = ctx.fitc_eval(theta, U, kind)
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 3, 1))
    try:
        b = ctx.fitc_eval(theta, U, kind)
    finally:
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 3, 33))
    assert abs(a[0] - b[0]) <= 1e-10 * abs(a[0])
    assert relerr(a[1], b[1]) <= 1e-8 and relerr(a[2], b[2]) <= 1e-8


def test_large_m_block_objective_after_loo_score_and_back(ctx):
    """Alternating score families on one context: the block path's scratch aliases the LOO vectors."""
    from oracle import woodbury as Wd
    from oracle import gp_oracle as O
    rng = np.random.default_rng(43)
    X = rng.uniform(-1, 1, (800, 4))
    y = np.sin(X @ rng.standard_normal(4)) + 0.1 * rng.standard_normal(800)
    U = rng.uniform(-1, 1, (48, 4))
    theta = np.concatenate([[0.1], np.zeros(4), [-2.0]])
    ctx.set_data(_dev(X), _dev(y))
    for kind in ("crps", "dss", "nlml", "kc", "logs"):
        val, grad, gU = ctx.fitc_eval(theta, U, kind)
        if kind in ("dss", "kc"):
            oval, og, ogU = Wd.fitc_block_obj_grad(X, y, U, theta, kind)
        else:
            oval, og, ogU = Wd.fitc_obj_grad(X, y, U, theta, O.SCORES[kind])[:3]
        assert abs(val - oval) <= OBJ_TOL * abs(oval), kind
        assert relerr(grad, og) <= GRAD_TOL and relerr(gU, ogU) <= GRAD_TOL, kind


@pytest.mark.parametrize("m_ind", [20, 48])
def test_block_objectives_through_sharded_entry_single_rank(m_ind):
    """gps_fitc_eval_sharded with a one-rank communicator: the row-offset exchange, the per-fold all-reduces and the
    replicated M x M chain of the block objectives, against gps_fitc_eval (row kernels at M = 20, matrix form at 48).
    The multi-rank case (folds straddling ranks) runs under torchrun in tests/mgpu_check.py."""
    import ctypes as C
    from gpscore_b200 import api, lib as L
    rng = np.random.default_rng(50 + m_ind)
    n = 1500                                        # 4 | n; fold size 375
    X = rng.uniform(-1, 1, (n, 8))
    y = np.sin(X @ rng.standard_normal(8)) + 0.1 * rng.standard_normal(n)
    U = rng.uniform(-1, 1, (m_ind, 8))
    theta = np.concatenate([[0.2], np.log(rng.uniform(0.8, 2.0, 8)), [-2.0]])
    c = api.Context(0)
    try:
        uid = (C.c_char * 128)()
        code = c._lib.gps_comm_unique_id(uid)
        if code != L.GPS_OK:
            pytest.skip("libnccl.so.2 not loadable")
        c._check(c._lib.gps_comm_init(c._h, C.c_char_p(uid.raw), 0, 1))
        c.set_data(_dev(X), _dev(y))
        for kind in ("dss", "kc"):
            a = c.fitc_eval(theta, U, kind)
            b = c.fitc_eval_sharded(theta, U, kind, n)
            assert abs(a[0] - b[0]) <= 1e-10 * abs(a[0]), kind
            assert relerr(b[1], a[1]) <= 1e-8 and relerr(b[2], a[2]) <= 1e-8, kind
        with pytest.raises(L.GpsError):             # the ranks' rows must add up to world_n
            c.fitc_eval_sharded(theta, U, "dss", n + 4)
    finally:
        c.close()


def test_large_m_rows_grow_inside_one_padded_shape(ctx):
    """N changes without changing the padded shape (1000 -> 1020 rows, both padded to 1024): the per-row scalar
    block is laid out with stride N and has to follow (it once kept the first N's size)."""
    from oracle import woodbury as Wd
    from oracle import gp_oracle as O
    rng = np.random.default_rng(77)
    X = rng.uniform(-1, 1, (1020, 8))
    y = np.sin(X @ rng.standard_normal(8)) + 0.1 * rng.standard_normal(1020)
    U = rng.uniform(-1, 1, (48, 8))
    theta = np.concatenate([[0.2], np.log(rng.uniform(0.8, 2.0, 8)), [-2.0]])
    for n in (1000, 1020):
        ctx.set_data(_dev(X[:n]), _dev(y[:n]))
        for kind in ("crps", "kc"):
            val, grad, gU = ctx.fitc_eval(theta, U, kind)
            if kind == "kc":
                oval, og, ogU = Wd.fitc_block_obj_grad(X[:n], y[:n], U, theta, kind)
            else:
                oval, og, ogU = Wd.fitc_obj_grad(X[:n], y[:n], U, theta, O.SCORES[kind])[:3]
            assert abs(val - oval) <= OBJ_TOL * abs(oval)
            assert relerr(grad, og) <= GRAD_TOL and relerr(gU, ogU) <= GRAD_TOL


@pytest.mark.parametrize("kind", ["crps", "dss"])
def test_large_m_failed_factorisation_is_reported_and_recoverable(ctx, kind):
    """No host synchronisation inside a matrix-form evaluation: a failed pivot is latched on the device and raised with
    the result read-back; the context is usable afterwards and refuses to predict from the failed evaluation."""
    from gpscore_b200 import lib as L
    from oracle import woodbury as Wd
    from oracle import gp_oracle as O
    rng = np.random.default_rng(91)
    X = rng.uniform(-1, 1, (800, 4))
    y = np.sin(X @ rng.standard_normal(4)) + 0.1 * rng.standard_normal(800)
    U = rng.uniform(-1, 1, (40, 4))
    theta = np.concatenate([[0.1], np.zeros(4), [-2.0]])
    ctx.set_data(_dev(X), _dev(y))
    with pytest.raises(L.GpsError) as ei:
        ctx.fitc_eval(theta, U, kind, jitter=-2.0)          # K_uu - 2 I is indefinite
    assert ei.value.code == L.GPS_ENOTPD and "K_uu" in str(ei.value)
    xs, mo, vo = _dev(X[:4]), _dev(np.zeros(4)), _dev(np.zeros(4))
    with pytest.raises(L.GpsError):
        ctx._check(ctx._lib.gps_fitc_predict(ctx._h, xs.data_ptr(), 4, mo.data_ptr(), vo.data_ptr()))
    val, grad, gU = ctx.fitc_eval(theta, U, kind)
    if kind == "dss":
        oval, og, ogU = Wd.fitc_block_obj_grad(X, y, U, theta, kind)
    else:
        oval, og, ogU = Wd.fitc_obj_grad(X, y, U, theta, O.SCORES[kind])[:3]
    assert abs(val - oval) <= OBJ_TOL * abs(oval)
    assert relerr(grad, og) <= GRAD_TOL and relerr(gU, ogU) <= GRAD_TOL
