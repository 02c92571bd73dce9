"""Round-2 parity tests (-m gpu), all through the C-ABI:

* the headline sizes against the CPU oracle itself (N = 10 000 full GP and FITC M = 20) — not only through
  finite differences;
* the fused FITC path: device-resident descent loop, objective-only evaluations, LOO read-back, prediction,
  ragged and tiny shapes, D up to 16;
* the same-signature twins Q / cal_mean_and_cov / spgp_cal_mean_and_cov (products on the library's GEMM);
* the advisor's findings: stream ordering against torch, training-set cache identity, staged protocol
  rejecting the block objectives.
Tolerances: objective 1e-8 relative, gradients 1e-6 relative (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, relerr

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


# ---- headline sizes against the oracle ---------------------------------------------------------------------------
def test_full_N10000_vs_oracle(ctx):
    """BASELINE's full-GP size: objective and gradient against oracle/gp_oracle.py (~10 s of CPU)."""
    from gpscore_b200 import synth
    from oracle import gp_oracle as O
    X, y = synth.kin40k_like(10000)
    theta = synth.hyper_point("P1")
    ctx.set_data(_dev(X), _dev(y))
    val, grad = ctx.full_eval(theta, "crps")
    oval, ograd = O.full_obj_grad(X, y, theta, O.SCORE_CRPS)
    assert abs(val - oval) <= OBJ_TOL * abs(oval), (val, oval)
    assert relerr(grad, ograd) <= GRAD_TOL


@pytest.mark.parametrize("score", ["crps", "logs", "nlml"])
def test_fitc20_N10000_vs_oracle(ctx, score):
    """BASELINE's FITC size (N = 10 000, M = 20) against the O(N M^2) CPU restatement, which is itself pinned
    to the dense reference goldens in tests/test_oracle.py."""
    from gpscore_b200 import synth
    from oracle import gp_oracle as O
    from oracle import woodbury as W
    X, y = synth.kin40k_like(10000)
    theta = synth.hyper_point("P1")
    U = synth.inducing_init(20)
    ctx.set_data(_dev(X), _dev(y))
    val, g, gU = ctx.fitc_eval(theta, U, score)
    oval, og, ogU, om, ov = W.fitc_obj_grad(X, y, U, theta, O.SCORES[score])
    assert abs(val - oval) <= OBJ_TOL * abs(oval), (val, oval)
    assert relerr(g, og) <= GRAD_TOL and relerr(gU, ogU) <= GRAD_TOL
    m, v = ctx.fitc_loo()
    assert relerr(m.cpu().numpy().ravel(), om) <= 1e-8 and relerr(v.cpu().numpy().ravel(), ov) <= 1e-8


def test_fitc20_N1e6_vs_oracle(ctx):
    """The scaling-sweep point N = 10^6, M = 20 against the CPU Woodbury restatement (~10 s of CPU)."""
    from gpscore_b200 import synth
    from oracle import gp_oracle as O
    from oracle import woodbury as W
    X, y = synth.kin40k_like(1000000, seed=7)
    theta = synth.hyper_point("P1")
    U = synth.inducing_init(20)
    ctx.set_data(_dev(X), _dev(y))
    val, g, gU = ctx.fitc_eval(theta, U, "crps")
    oval, og, ogU = W.fitc_obj_grad(X, y, U, theta, O.SCORE_CRPS)[:3]
    assert abs(val - oval) <= OBJ_TOL * abs(oval), (val, oval)
    assert relerr(g, og) <= GRAD_TOL and relerr(gU, ogU) <= GRAD_TOL
    ctx.set_data(_dev(X[:64]), _dev(y[:64]))     # release the large workspaces' data


# ---- fused FITC path ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,m_ind,d", [(1, 1, 1), (7, 3, 2), (33, 7, 1), (257, 8, 3), (1000, 15, 8), (999, 16, 9),
                                        (513, 23, 16), (2049, 24, 5), (640, 31, 8), (31, 20, 8)])
def test_fused_shapes_vs_oracle(ctx, n, m_ind, d):
    """Every operand-shape instantiation of the fused kernels (M / 8 + 1 row tiles, ceil(M / 4) k-steps, D <= 8 and
    D <= 16), ragged row counts below and above the 32-row warp tile, against the dense oracle."""
    from oracle import gp_oracle as O
    rng = np.random.default_rng(1000 + n + m_ind)
    X = rng.standard_normal((n, d))
    y = rng.standard_normal((n, 1))
    U = rng.standard_normal((m_ind, d))
    theta = np.concatenate([[0.3], rng.uniform(0.0, 1.0, d), [-0.7]])
    ctx.set_data(_dev(X), _dev(y))
    for score in ("crps", "logs", "nlml"):
        val, g, gU = ctx.fitc_eval(theta, U, score)
        oval, og, ogU, _ = O.fitc_obj_grad(X, y, U, theta, O.SCORES[score])
        assert abs(val - oval) <= OBJ_TOL * abs(oval), (score, val, oval)
        # (n = 1 has identically zero length-scale / inducing-input gradients: absolute floor for those)
        assert np.max(np.abs(g - og)) <= GRAD_TOL * max(np.max(np.abs(og)), 1e-6), score
        assert np.max(np.abs(gU - ogU)) <= GRAD_TOL * max(np.max(np.abs(ogU)), 1e-6), score
    Xs = rng.standard_normal((45, d))
    mean, var = ctx.fitc_predict(theta, U, _dev(Xs))
    om, ov = O.fitc_predict(X, y, U, Xs, theta)
    assert relerr(mean.cpu().numpy(), om) <= 1e-8 and relerr(var.cpu().numpy(), ov) <= 1e-7


def test_fused_objective_only_matches(ctx):
    from gpscore_b200 import api as A, lib as L, synth
    X, y = synth.kin40k_like(1500, seed=3)
    theta = synth.hyper_point("P2")
    U = synth.inducing_init(20, seed=4)
    ctx.set_data(_dev(X), _dev(y))
    th, Uh = np.ascontiguousarray(theta), np.ascontiguousarray(U).ravel()
    for score in ("crps", "logs", "nlml"):
        ref = ctx.fitc_eval(theta, U, score)[0]
        obj = np.zeros(1)
        ctx._check(ctx._lib.gps_fitc_eval(ctx._h, A._dp(th), A._dp(Uh), 20, 1e-3, L.SCORES[score], A._dp(obj), None, None))
        assert abs(obj[0] - ref) <= 1e-13 * abs(ref), score


@pytest.mark.parametrize("score,lr,lr_u", [("crps", 0.5, 0.5), ("nlml", 1e-4, 1e-5), ("logs", 0.3, 0.1)])
def test_fused_descend_equals_host_loop(ctx, score, lr, lr_u):
    """gps_fitc_descend (theta and U resident on the device, K20:243-251 with the two learning rates of
    K20:326-327) reproduces the loop driven from the host one evaluation at a time."""
    from gpscore_b200 import synth
    X, y = synth.kin40k_like(800, seed=11)
    theta = synth.hyper_point("P1")
    U = synth.inducing_init(20, seed=12)
    ctx.set_data(_dev(X), _dev(y))
    iters = 12
    th, Uc, trace = theta.copy(), U.copy(), []
    for _ in range(iters):
        v, g, gU = ctx.fitc_eval(th, Uc, score)
        trace.append(v)
        th = th - lr * g
        Uc = Uc - lr_u * gU
    th2, U2, tr2 = ctx.fitc_descend(theta, U, score, lr, lr_u, iters)
    assert relerr(tr2, np.array(trace)) <= 1e-12
    assert relerr(th2, th) <= 1e-12 and relerr(U2, Uc) <= 1e-12


def test_fused_descend_reports_failed_factorisation(ctx):
    from gpscore_b200 import lib as L
    rng = np.random.default_rng(5)
    X = rng.standard_normal((200, 2))
    y = rng.standard_normal((200, 1))
    ctx.set_data(_dev(X), _dev(y))
    U = np.zeros((6, 2))                                  # identical inducing points and a NEGATIVE jitter: not PD
    with pytest.raises(L.NotPositiveDefinite):
        ctx.fitc_descend(np.zeros(4), U, "crps", 0.1, 0.1, 3, jitter=-1.0)
    with pytest.raises(L.NotPositiveDefinite):
        ctx.fitc_eval(np.zeros(4), U, "crps", jitter=-1.0)
    # the context stays usable
    val, _, _ = ctx.fitc_eval(np.zeros(4), rng.standard_normal((6, 2)), "crps")
    assert np.isfinite(val)


def test_fused_is_deterministic(ctx):
    from gpscore_b200 import synth
    X, y = synth.kin40k_like(50000, seed=21)
    theta = synth.hyper_point("P1")
    U = synth.inducing_init(20, seed=22)
    ctx.set_data(_dev(X), _dev(y))
    a = ctx.fitc_eval(theta, U, "crps")
    for _ in range(3):
        b = ctx.fitc_eval(theta, U, "crps")
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


# ---- same-signature twins -----------------------------------------------------------------------------------------------
def test_twins_Q_and_predictive_moments_vs_golden(ctx):
    """Q (KF:32-39), cal_mean_and_cov (KF:121-126) and spgp_cal_mean_and_cov (K20:76-83) with the reference's
    positional signatures reproduce the golden predictions made by the reference's own code."""
    from gpscore_b200 import api
    from oracle import gp_oracle as O
    g = load_golden("c3_kin_full_P1")
    th = g["theta"]
    pk, pl, pn = torch.tensor([th[0]]), torch.tensor(th[1:-1].reshape(1, -1)), torch.tensor([th[-1]])
    X, y, Xs = _dev(g["X"]), _dev(g["y"]), _dev(g["Xs"])
    n, t = X.shape[0], Xs.shape[0]
    api.sigma_noise_sq = torch.exp(pn)
    k1, k2, k3 = api.ARD(Xs, X, pk, pl), api.ARD(X, X, pk, pl), api.ARD(Xs, Xs, pk, pl)
    mean, cov = api.cal_mean_and_cov(k1, k2, k3, t, n, y)
    assert relerr(mean.cpu().numpy(), g["pred_mean"]) <= 1e-8
    assert relerr(torch.diag(cov).cpu().numpy(), g["pred_var"].ravel()) <= 1e-8
    gf = load_golden("c4_kin_fitc_P1")
    th = gf["theta"]
    pk, pl, pn = torch.tensor([th[0]]), torch.tensor(th[1:-1].reshape(1, -1)), torch.tensor([th[-1]])
    X, y, Xs, U = _dev(gf["X"]), _dev(gf["y"]), _dev(gf["Xs"]), _dev(gf["U"])
    n, t = X.shape[0], Xs.shape[0]
    api.para_k, api.para_l, api.sigma_noise_sq = pk, pl, torch.exp(pn)
    Qff = api.Q(X, U, X)
    Qref = O.Q(gf["X"], gf["U"], gf["X"], th[0], th[1:-1])
    assert relerr(Qff.cpu().numpy(), Qref) <= 1e-9
    Qsf = api.Q(Xs, U, X)
    kff, kss = api.ARD(X, X, pk, pl), api.ARD(Xs, Xs, pk, pl)
    mean, cov = api.spgp_cal_mean_and_cov(kff, Qff, Qsf, kss, t, n, y)
    assert relerr(mean.cpu().numpy(), gf["pred_mean"]) <= 1e-7
    assert relerr(torch.diag(cov).cpu().numpy(), gf["pred_var"].ravel()) <= 1e-7


def test_matmul_twin_ragged(ctx):
    rng = np.random.default_rng(3)
    for m, k, n in ((1, 1, 1), (130, 257, 5), (300, 128, 129)):
        A, B = rng.standard_normal((m, k)), rng.standard_normal((k, n))
        out = ctx.matmul(_dev(A), _dev(B)).cpu().numpy()
        assert relerr(out, A @ B) <= 1e-13
    assert relerr(ctx.matmul(torch.from_numpy(A), torch.from_numpy(B)).numpy(), A @ B) <= 1e-13   # host buffers


# ---- advisor findings ---------------------------------------------------------------------------------------------------
def test_staged_protocol_rejects_block_objectives(ctx):
    from gpscore_b200 import lib as L, synth
    X, y = synth.kin40k_like(400, seed=2)
    ctx.set_data(_dev(X), _dev(y))
    for kind in ("dss", "kc"):
        with pytest.raises(L.GpsError):
            ctx.fitc_eval_sharded(synth.hyper_point("P1"), synth.inducing_init(20), kind, 400, lambda t: t)
    # ... while the fused single-GPU entry point still evaluates them
    assert np.isfinite(ctx.fitc_eval(synth.hyper_point("P1"), synth.inducing_init(20), "dss")[0])


def test_library_is_ordered_behind_torch_streams(ctx):
    """Inputs produced on a side stream right before the call (a long chain of torch ops ending in the
    training set) are seen complete: the context follows torch's current stream."""
    from gpscore_b200 import api, synth
    from oracle import gp_oracle as O
    X, y = synth.kin40k_like(2000, seed=8)
    theta = synth.hyper_point("P1")
    c = api.Context(0)
    side = torch.cuda.Stream()
    base = _dev(X)
    with torch.cuda.stream(side):
        Xd = base.clone()
        for _ in range(200):                               # keep the stream busy before the data is final
            Xd = Xd + 1.0
        Xd = Xd - 200.0
        yd = _dev(y) * 1.0
        c.set_data(Xd, yd)
        val, grad = c.full_eval(theta, "crps")
    oval, og = O.full_obj_grad(X, y, theta, O.SCORE_CRPS)
    assert abs(val - oval) <= 1e-8 * abs(oval) and relerr(grad, og) <= 1e-6
    c.close()


def test_training_set_cache_is_keyed_on_identity(ctx):
    """full_loo_objective re-uploads when a NEW tensor (possibly recycled at the same address by the caching
    allocator) is passed, and after a direct set_data."""
    from gpscore_b200 import api, synth
    from oracle import gp_oracle as O
    c = api.Context(0)
    theta = synth.hyper_point("P1")
    pk, pl, pn = (torch.tensor([theta[0]], requires_grad=True), torch.tensor(theta[1:-1].reshape(1, -1), requires_grad=True),
                  torch.tensor([theta[-1]], requires_grad=True))
    vals = []
    for seed in (1, 2, 3):
        X, y = synth.kin40k_like(256, seed=seed)
        Xd, yd = _dev(X), _dev(y)                          # freed at the end of the iteration: the address recycles
        v = api.full_loo_objective(Xd, yd, pk, pl, pn, "crps", ctx=c)
        ov, _ = O.full_obj_grad(X, y, theta, O.SCORE_CRPS)
        assert abs(float(v) - ov) <= 1e-8 * abs(ov), seed
        vals.append(float(v))
        del Xd, yd
    assert len(set(vals)) == 3
    X1, y1 = synth.kin40k_like(256, seed=1)
    X2, y2 = synth.kin40k_like(256, seed=2)
    X1d, y1d = _dev(X1), _dev(y1)
    v1 = float(api.full_loo_objective(X1d, y1d, pk, pl, pn, "crps", ctx=c))
    c.set_data(_dev(X2), _dev(y2))                         # direct upload invalidates the cached key
    v1b = float(api.full_loo_objective(X1d, y1d, pk, pl, pn, "crps", ctx=c))
    assert v1 == v1b
    c.close()


def test_multi_gpu_parity_under_torchrun():
    """tests/mgpu_check.py on two GPUs (library NCCL all-reduces inside the row-sharded evaluation, sharded
    prediction, sharded grid).  Skipped on a one-GPU box; the driver's scaling run covers it through bench.py's
    own sharded-vs-single parity gates."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


# ---- the experiment drivers ----------------------------------------------------------------------------------
def _run_example(args, timeout=600):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable] + args, cwd=root, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_example_kin40k_compare_runs_both_models():
    """examples/kin40k_compare.py (the trial loop KF:190-299 / K20:184-304 on the fused objectives): both models run,
    the host-driven loop and the device-resident loop end at the same objective, and the result table is printed."""
    import re
    out_h = _run_example(["examples/kin40k_compare.py", "--model", "fitc", "--trials", "1", "--n-train", "300", "--n-test", "200",
                          "--itr-scale", "0.01"])
    out_d = _run_example(["examples/kin40k_compare.py", "--model", "fitc", "--trials", "1", "--n-train", "300", "--n-test", "200",
                          "--itr-scale", "0.01", "--on-device"])
    assert "means over trials" in out_h and "fitted by crps" in out_h and "fitted by kc" in out_h

    def last_obj(out, score):
        return float(re.search(r"trial 0 %s\s+itr\s+\d+ objective ([-0-9.e+]+)" % score, out).group(1))

    for score in ("crps", "nlml", "logs"):
        # the printed objective is the one BEFORE the last update in both variants
        assert abs(last_obj(out_h, score) - last_obj(out_d, score)) <= 1e-5 * abs(last_obj(out_h, score)), score
    out_f = _run_example(["examples/kin40k_compare.py", "--model", "full", "--trials", "1", "--n-train", "256", "--n-test", "100",
                          "--itr-scale", "0.02"])
    assert "fitted by dss" in out_f


def test_example_contour_grid_writes_reference_layout(tmp_path):
    """examples/contour_grid.py: the four 50 x 50 matrices of CP:113-141 (rows = noise s.d., columns = length scale),
    spot-checked against the R twins in the oracle."""
    from oracle import gp_oracle as O
    out = str(tmp_path / "contour")
    _run_example(["examples/contour_grid.py", "--n", "20", "--grid", "50", "--out", out])
    rng = np.random.default_rng(0)
    x = np.linspace(-6, 6, 20)
    K = np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2) + 1e-10 * np.eye(20)
    y = (np.linalg.cholesky(K) @ rng.standard_normal(20) + 0.1 * rng.standard_normal(20)).reshape(-1, 1)
    l_range, noise_range = np.linspace(0.01, 2, 50), np.linspace(0.01, 1, 50)
    for which, fn in (("crps", O.cal_m_crps), ("nlml", O.cal_NLML), ("logs", O.cal_m_logs), ("wrong_crps", O.wrong_cal_m_crps)):
        mat = np.load("%s/ma_%s.npy" % (out, which))
        assert mat.shape == (50, 50)
        csv = np.loadtxt("%s/ma_%s.csv" % (out, which), delimiter=",")
        assert np.allclose(csv, mat, rtol=1e-12, atol=0)
        for (i, j) in ((5, 7), (20, 30), (49, 49), (12, 3)):
            want = fn(x, y, l_range[j], noise_range[i])
            assert abs(mat[i, j] - want) <= 1e-8 * max(1.0, abs(want)), (which, i, j)


@pytest.mark.parametrize("score,lr", [("crps", 0.5), ("logs", 0.05), ("nlml", 5e-4)])
def test_full_descend_device_resident_equals_host_loop(ctx, score, lr):
    """gps_full_descend keeps theta on the device for crps / logs / nlml (parameter kernel, evaluation, update kernel
    enqueued back to back, KF:237-260): identical trajectory to the loop driven one evaluation at a time."""
    from gpscore_b200 import synth
    X, y = synth.kin40k_like(700, seed=31)
    theta = synth.hyper_point("P1")
    ctx.set_data(_dev(X), _dev(y))
    th, trace = theta.copy(), []
    for _ in range(8):
        v, g = ctx.full_eval(th, score)
        trace.append(v)
        th = th - lr * g
    th2, tr2 = ctx.full_descend(theta, score, lr, 8)
    assert relerr(tr2, np.array(trace)) <= 1e-12 and relerr(th2, th) <= 1e-12


def test_full_descend_reports_failed_factorisation(ctx):
    from gpscore_b200 import lib as L
    rng = np.random.default_rng(2)
    X = np.repeat(rng.standard_normal((40, 2)), 5, axis=0)          # duplicated rows ...
    y = rng.standard_normal((200, 1))
    ctx.set_data(_dev(X), _dev(y))
    theta = np.array([0.0, 0.0, 0.0, -800.0])                       # ... and no noise: K is singular
    with pytest.raises(L.NotPositiveDefinite):
        ctx.full_descend(theta, "crps", 0.1, 3)
    assert np.isfinite(ctx.full_eval(np.array([0.0, 0.0, 0.0, 0.0]), "crps")[0])
