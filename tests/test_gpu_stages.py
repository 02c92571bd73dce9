"""Stage-level GPU tests (-m gpu): each dense kernel against numpy on the same inputs.
fp64 tolerances are written next to each assert."""
import ctypes as C

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


@pytest.fixture
def restore_variants(ctx):
    yield
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 0, 6))
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 1, 1))


@pytest.mark.parametrize("variant", [0, 2, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("shape", [(128, 128, 128), (256, 384, 512)])
def test_tile_gemm(ctx, restore_variants, kind, shape, variant):
    """every tile policy (CTA tile, stage depth, warp grid) gives the same product"""
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 0, variant))
    Mp, Np, Kp = shape
    rng = np.random.default_rng(kind * 10 + Mp)
    A = rng.standard_normal((Mp, Kp))
    B = rng.standard_normal((Np, Kp))
    C0 = rng.standard_normal((Mp, Np))
    ref = 0.5 * A @ B.T - 2.0 * C0
    Ad = _dev(A if kind != 2 else A.T.copy())
    Bd = _dev(B if kind == 0 else B.T.copy())
    Cd = _dev(C0)
    ctx._check(ctx._lib.gps_dbg_gemm(ctx._h, kind, Ad.data_ptr(), Bd.data_ptr(), Cd.data_ptr(), Mp, Np, Kp,
                                     0.5, -2.0, None, 0))
    assert relerr(Cd.cpu().numpy(), ref) < 1e-13


@pytest.mark.parametrize("variant", [0, 5, 6, 8])
def test_tile_gemm_dvec_and_mirror(ctx, restore_variants, variant):
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 0, variant))
    n = 384
    rng = np.random.default_rng(3)
    A = rng.standard_normal((n, n))
    dv = rng.standard_normal(n)
    Cd = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    ctx._check(ctx._lib.gps_dbg_gemm(ctx._h, 0, _dev(A).data_ptr(), _dev(A).data_ptr(), Cd.data_ptr(), n, n, n,
                                     1.0, 0.0, _dev(dv).data_ptr(), 0))
    assert relerr(Cd.cpu().numpy(), (A * dv) @ A.T) < 1e-13
    Cd.zero_()
    ctx._check(ctx._lib.gps_dbg_gemm(ctx._h, 2, _dev(A).data_ptr(), _dev(A).data_ptr(), Cd.data_ptr(), n, n, n,
                                     1.0, 0.0, None, 1))
    assert relerr(Cd.cpu().numpy(), A.T @ A) < 1e-13   # lower tiles computed, upper tiles mirrored


@pytest.mark.parametrize("diag_kernel", [0, 1])
@pytest.mark.parametrize("n", [1, 50, 128, 129, 300, 640, 1000])
def test_factor_inverse(ctx, restore_variants, n, diag_kernel):
    """both diagonal-block kernels (register-cyclic, 32-blocked DMMA) through POTRF/TRTRI/LAUUM"""
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 1, diag_kernel))
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n + 5))
    A = G @ G.T / n + 0.5 * np.eye(n)
    L = torch.empty(n, n, dtype=torch.float64, device="cuda")
    Li = torch.empty_like(L)
    Ai = torch.empty_like(L)
    ctx._check(ctx._lib.gps_dbg_factor(ctx._h, _dev(A).data_ptr(), n, L.data_ptr(), Li.data_ptr(), Ai.data_ptr()))
    Lr = np.linalg.cholesky(A)
    assert relerr(np.tril(L.cpu().numpy()), Lr) < 1e-12
    assert relerr(Li.cpu().numpy(), np.linalg.inv(Lr)) < 1e-11
    assert relerr(Ai.cpu().numpy(), np.linalg.inv(A)) < 1e-11


def test_not_positive_definite_is_reported(ctx):
    from gpscore_b200 import lib as L
    n = 200
    A = -np.eye(n)
    out = torch.empty(n, n, dtype=torch.float64, device="cuda")
    with pytest.raises(L.NotPositiveDefinite):
        ctx._check(ctx._lib.gps_dbg_factor(ctx._h, _dev(A).data_ptr(), n, out.data_ptr(), None, None))
    with pytest.raises(RuntimeError):   # the reference convention: a failed Cholesky is a RuntimeError
        ctx._check(ctx._lib.gps_dbg_factor(ctx._h, _dev(A).data_ptr(), n, out.data_ptr(), None, None))


@pytest.mark.parametrize("n,d", [(1, 1), (120, 1), (333, 8), (500, 8), (200, 13)])
def test_gram_and_ard(ctx, n, d):
    from oracle import gp_oracle as O
    rng = np.random.default_rng(n + d)
    X = rng.standard_normal((n, d))
    y = rng.standard_normal(n)
    theta = np.concatenate([[0.3], rng.random(d), [-1.0]])
    ctx.set_data(_dev(X), _dev(y))
    K = torch.empty(n, n, dtype=torch.float64, device="cuda")
    th = np.ascontiguousarray(theta)
    ctx._check(ctx._lib.gps_dbg_gram(ctx._h, th.ctypes.data_as(C.POINTER(C.c_double)), K.data_ptr()))
    ref = O.ARD(X, X, theta[0], theta[1:-1]) + np.exp(theta[-1]) * np.eye(n)
    assert relerr(np.tril(K.cpu().numpy()), np.tril(ref)) < 1e-13
    # rectangular twin, device and host inputs, 1-element (broadcast) and full b (KF:8)
    Xp = rng.standard_normal((77, d))
    for b in (theta[1:-1], theta[1:2]):
        out = ctx.ard(_dev(X), _dev(Xp), theta[0], b)
        assert relerr(out.cpu().numpy(), O.ARD(X, Xp, theta[0], b)) < 1e-13
        out = ctx.ard(torch.from_numpy(X), torch.from_numpy(Xp), theta[0], b)   # host pointers
        assert out.device.type == "cpu"
        assert relerr(out.numpy(), O.ARD(X, Xp, theta[0], b)) < 1e-13


def test_fp64_peak_probe(ctx):
    a, b = C.c_double(), C.c_double()
    ctx._check(ctx._lib.gps_dbg_fp64_peak(ctx._h, 2000, C.byref(a), C.byref(b)))
    print("DMMA %.2f TFLOP/s  DFMA %.2f TFLOP/s" % (a.value, b.value))
    assert a.value > 1.0 and b.value > 1.0
