"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in dist.py: row partition, the
three-all-reduce FITC protocol, sharded metric sums and the round-robin grid sweep.  The compute
backend here is a numpy stand-in with the SAME staged protocol as the CUDA passes (built from
oracle/woodbury.py pieces); on the GPU box the identical host code drives api.Context over NCCL."""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden, grad_vector, relerr


def staged(X, y, U, theta, score, world_n, allreduce):
    """Per-rank FITC passes on this rank's rows; `allreduce` sums a flat float64 tensor in place at
    exactly the three points where the CUDA path hands its packed accumulators to NCCL."""
    from scipy.linalg import cholesky, solve_triangular
    from oracle import gp_oracle as O
    from oracle.woodbury import _kern, _phi_adj
    a, b, c = O._split(theta)
    n, D = X.shape
    m = U.shape[0]
    ell = np.exp(np.asarray(b, dtype=np.float64).ravel())
    if ell.size == 1:
        ell = np.full(D, ell[0])
    ea, sn2 = math.exp(a), math.exp(c)
    y = y.reshape(-1)

    def red(*arrs):
        flat = torch.from_numpy(np.concatenate([np.asarray(x, dtype=np.float64).ravel() for x in arrs]))
        allreduce(flat)
        out, o, f = [], 0, flat.numpy()
        for x in arrs:
            k = np.asarray(x).size
            out.append(f[o:o + k].reshape(np.asarray(x).shape).copy())
            o += k
        return out

    Kuu = _kern(U, U, a, ell)
    LA = cholesky(Kuu + O.JITTER * np.eye(m), lower=True)
    Kuf = _kern(U, X, a, ell)
    V = solve_triangular(LA, Kuf, lower=True)
    lam = ea - np.sum(V * V, axis=0) + sn2
    Cm, vy = red((V / lam) @ V.T, V @ (y / lam))                               # all-reduce 1
    LC = cholesky(np.eye(m) + Cm, lower=True)
    beta = solve_triangular(LC, vy, lower=True)
    W = solve_triangular(LC, V, lower=True)
    r = np.sum(W * W, axis=0)
    d = 1.0 / lam - r / lam ** 2
    alpha = (y - W.T @ beta) / lam
    if score == O.SCORE_NLML:
        obj = 0.5 * np.sum(np.log(lam)) + 0.5 * float(y @ alpha)
        abar, dbar, lam_bar = 0.5 * y, np.zeros_like(d), 0.5 / lam
    else:
        v, abar, dbar = O._score_and_seeds(alpha.reshape(-1, 1), d.reshape(-1, 1), score)
        k = n / world_n
        obj, abar, dbar, lam_bar = v * k, abar.ravel() * k, dbar.ravel() * k, np.zeros_like(lam)
    lam_bar = lam_bar + dbar * (-1.0 / lam ** 2 + 2.0 * r / lam ** 3) - abar * alpha / lam
    rbar, tbar = -dbar / lam ** 2, -abar / lam
    R, beta_bar, objv = red((W * rbar) @ W.T, W @ tbar, np.array([obj]))        # all-reduce 2
    obj = float(objv[0])
    if score == O.SCORE_NLML:
        obj += 0.5 * world_n * math.log(2 * math.pi) + np.sum(np.log(np.diag(LC)))
        LC_bar0 = np.diag(1.0 / np.diag(LC))
    else:
        LC_bar0 = np.zeros((m, m))
    SW = np.outer(beta, beta_bar) + 2.0 * R + np.outer(beta_bar, beta)
    C_bar = _phi_adj(LC, -np.tril(solve_triangular(LC, SW, lower=True, trans="T")) + LC_bar0)
    vy_bar = solve_triangular(LC, beta_bar, lower=True, trans="T")
    CV = C_bar @ V
    lam_bar = lam_bar - (beta_bar @ W) * y / lam ** 2 - np.sum(V * CV, axis=0) / lam ** 2
    Vbar = solve_triangular(LC, np.outer(beta, tbar) + 2.0 * W * rbar, lower=True, trans="T") \
        + np.outer(vy_bar, y / lam) + 2.0 * CV / lam - 2.0 * V * lam_bar
    G = solve_triangular(LA, Vbar, lower=True, trans="T") * Kuf
    gb = np.zeros(D)
    gU = np.zeros((m, D))
    for dd in range(D):
        diff = U[:, dd][:, None] - X[:, dd][None, :]
        gb[dd] = np.sum(G * diff * diff) / ell[dd] ** 2
        gU[:, dd] = -np.sum(G * diff, axis=1) / ell[dd] ** 2
    S, slb, ga, gb, gU = red(Vbar @ V.T, np.array([lam_bar.sum()]), np.array([G.sum()]), gb, gU)  # all-reduce 3
    A_bar = _phi_adj(LA, -np.tril(solve_triangular(LA, S, lower=True, trans="T")))
    G2 = A_bar * Kuu
    g_a = ea * slb[0] + ga[0] + G2.sum()
    g_c = sn2 * slb[0]
    for dd in range(D):
        diff = U[:, dd][:, None] - U[:, dd][None, :]
        gb[dd] += np.sum(G2 * diff * diff) / ell[dd] ** 2
        gU[:, dd] += -2.0 * np.sum(G2 * diff, axis=1) / ell[dd] ** 2
    return obj, np.concatenate([[g_a], gb, [g_c]]), gU


class NumpyStagedFitc:
    """Stand-in for api.Context.fitc_eval_sharded (same signature)."""

    def __init__(self, X, y):
        self.X, self.y = X, y

    def fitc_eval_sharded(self, theta, U, score, world_n, allreduce):
        from oracle import gp_oracle as O
        return staged(self.X, self.y, U, theta, O.SCORES[score], world_n, allreduce)


def _worker(rank, size, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from scipy.special import erf
        from gpscore_b200 import dist as D
        from oracle import gp_oracle as O
        g = load_golden("c4_kin_fitc_ragged")           # N = 333: uneven split
        n = g["X"].shape[0]
        lo, hi = D.row_block(n, rank, size)
        fitc = D.ShardedFitc(NumpyStagedFitc(g["X"][lo:hi], g["y"][lo:hi]), n)
        res = {}
        for score in ("crps", "logs", "nlml"):
            val, grad, gU = fitc.eval(g["theta"], g["U"], score)
            res[score] = (abs(val - g["obj_" + score]) / abs(g["obj_" + score]),
                          relerr(grad, grad_vector(g, score)), relerr(gU, g["grad_u_" + score]))
        # sharded test metrics: rows of the test set
        mean, var = O.fitc_predict(g["X"], g["y"], g["U"], g["Xs"], g["theta"])
        t = g["Xs"].shape[0]
        tlo, thi = D.row_block(t, rank, size)
        m, c, yy = mean[tlo:thi].ravel(), var[tlo:thi].ravel(), g["ys"][tlo:thi].ravel()
        ytm, ytv = g["y"].mean(), g["y"].var(ddof=1)
        sd = np.sqrt(c)
        z = (yy - m) / sd
        crps_i = sd * (z * erf(z / np.sqrt(2)) + 2 * np.exp(-0.5 * z * z) / np.sqrt(2 * np.pi) - 1 / np.sqrt(np.pi))
        logs_i = (yy - m) ** 2 / (2 * c) + 0.5 * np.log(c) + 0.5 * np.log(2 * np.pi)
        triv_i = 0.5 * np.log(2 * np.pi * ytv) + (yy - ytm) ** 2 / (2 * ytv)
        inside = ((m + 2 * sd - yy) > 0) & ((yy - (m - 2 * sd)) > 0)
        sums = [((yy - m) ** 2).sum(), ((ytm - yy) ** 2).sum(), logs_i.sum(), crps_i.sum(), triv_i.sum(),
                float(inside.sum())]
        met = D.sharded_metrics(sums, t)
        res["metrics"] = max(abs(met[k] - float(g["m_" + k])) for k in met)
        # round-robin grid sweep: CP:109-144 on the 1-D toy problem
        x = np.linspace(-6, 6, 20)
        yv = np.sin(x).reshape(-1, 1)
        Lg, Sg = np.meshgrid(np.linspace(0.2, 2, 5), np.linspace(0.05, 1, 4), indexing="ij")
        fn = lambda ls, sd: [O.cal_m_crps(x, yv, l, s) for l, s in zip(ls, sd)]
        got = D.sharded_grid(fn, Lg.ravel(), Sg.ravel())
        want = np.array(fn(Lg.ravel(), Sg.ravel()))
        res["grid"] = float(np.max(np.abs(got - want)))
        res["grid_shape"] = D.grid_matrix(got, 5, 4).shape
        ret[rank] = res
    finally:
        dist.destroy_process_group()


def test_row_block_and_round_robin():
    from gpscore_b200 import dist as D
    for n in (1, 7, 333, 10000):
        for size in (1, 2, 3, 8):
            blocks = [D.row_block(n, r, size) for r in range(size)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(size - 1))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
            idx = np.sort(np.concatenate([D.round_robin(n, r, size) for r in range(size)]))
            assert np.array_equal(idx, np.arange(n))


def test_world_size_2_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert len(ret) == 2
    for rank in range(2):
        r = ret[rank]
        for score in ("crps", "logs", "nlml"):
            assert r[score][0] <= 1e-8 and r[score][1] <= 1e-6 and r[score][2] <= 1e-6, (rank, score, r[score])
        assert r["metrics"] <= 1e-7
        assert r["grid"] <= 1e-14
        assert r["grid_shape"] == (4, 5)
