"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in dist.py: row partition, the
three-all-reduce FITC protocol, the row-sharded block-objective protocol (fold accumulators), sharded metric
sums and the round-robin grid sweep.  The compute
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


def staged_block(X, y, U, theta, kind, world_n, row_offset, allreduce):
    """The row-sharded 4-fold block objectives (DSS K20:538-587, kc K20:669-720) exactly as
    csrc/gps_fitc_large.cu::block_pass2 stages them: this rank holds the global rows
    [row_offset, row_offset + n); fold f is the global range [f nf, (f + 1) nf), nf = world_n / 4, of which a
    possibly empty part lives here.  All-reduce points: [C - I | v_y], the four folds' [P_f | g_f], the objective's row
    share, (kc) the four folds' [E_f | hbar_f], [S | kernel-gradient block | sum lam_bar].  G_W and beta_bar are
    functions of all-reduced matrices and therefore replicated."""
    from scipy.linalg import cholesky, solve_triangular
    from oracle import gp_oracle as O
    from oracle.woodbury import _kern, _phi_adj
    a, b, c = O._split(theta)
    n, D = X.shape
    m = U.shape[0]
    nf = world_n // 4
    ell = np.exp(np.asarray(b, dtype=np.float64).ravel())
    if ell.size == 1:
        ell = np.full(D, ell[0])
    ea, sn2 = math.exp(a), math.exp(c)
    y = y.reshape(-1)
    I = np.eye(m)

    def red(*arrs):
        flat = torch.from_numpy(np.concatenate([np.asarray(x, dtype=np.float64).ravel() for x in arrs]))
        allreduce(flat)
        out, o, f = [], 0, flat.numpy()
        for x in arrs:
            k = np.asarray(x).size
            out.append(f[o:o + k].reshape(np.asarray(x).shape).copy())
            o += k
        return out

    folds = [slice(min(n, max(0, f * nf - row_offset)), min(n, max(0, (f + 1) * nf - row_offset))) for f in range(4)]
    Kuu = _kern(U, U, a, ell)
    LA = cholesky(Kuu + O.JITTER * I, lower=True)
    Kuf = _kern(U, X, a, ell)
    V = solve_triangular(LA, Kuf, lower=True)
    lam = ea - np.sum(V * V, axis=0) + sn2
    Cm, vy = red((V / lam) @ V.T, V @ (y / lam))                               # all-reduce 1
    LC = cholesky(I + Cm, lower=True)
    beta = solve_triangular(LC, vy, lower=True)
    W = solve_triangular(LC, V, lower=True)
    alpha = (y - W.T @ beta) / lam
    # stage A: fold accumulators over this rank's rows
    P = np.stack([(W[:, sl] / lam[sl]) @ W[:, sl].T for sl in folds])
    gv = np.stack([W[:, sl] @ alpha[sl] for sl in folds])
    P, gv = red(P, gv)                                                         # all-reduce A
    abar, lam_bar, Dm = np.zeros(n), np.zeros(n), np.zeros((m, n))
    mbar, cbar = np.zeros(n), np.zeros(n)
    GW, beta_bar = np.zeros((m, m)), np.zeros(m)
    rows_obj, fold_obj = 0.0, 0.0
    Hinv, hv = [], []
    for f, sl in enumerate(folds):                                             # stage B
        LH = cholesky(I - P[f], lower=True)
        Hi = solve_triangular(LH, solve_triangular(LH, I, lower=True), lower=True, trans="T")
        h = Hi @ gv[f]
        Hinv.append(Hi)
        hv.append(h)
        Wf, lf, af = W[:, sl], lam[sl], alpha[sl]
        if kind == "dss":
            fold_obj += 0.5 * nf * math.log(2 * math.pi) - np.sum(np.log(np.diag(LH))) + 0.5 * float(gv[f] @ h)
            rows_obj += 0.5 * np.sum(np.log(lf)) + 0.5 * np.sum(lf * af * af)
            Hhat = -0.5 * Hi - 0.5 * np.outer(h, h)
            abar[sl] = lf * af + Wf.T @ h
            lam_bar[sl] = 0.5 / lf + 0.5 * af * af + np.sum(Wf * (Hhat @ Wf), axis=0) / lf ** 2
            Dm[:, sl] = -2.0 * (Hhat @ Wf) / lf + np.outer(h, af)
            GW += -2.0 * Hhat @ P[f] + np.outer(h, gv[f])
            beta_bar += -h
        else:
            HiW = Hi @ Wf
            sd = np.sqrt(lf + np.sum(Wf * HiW, axis=0))
            z = (lf * af + Wf.T @ h) / sd
            tpm1 = 2 * O._Phi(z) - 1
            rows_obj += float(np.sum(sd * (z * tpm1 + 2 * O._phi(z) - 1 / math.sqrt(math.pi)))) / nf
            mbar[sl] = -tpm1 / nf
            cbar[sl] = (2 * O._phi(z) - 1 / math.sqrt(math.pi)) / (2 * sd) / nf
            Dm[:, sl] = -np.outer(h, mbar[sl]) + 2.0 * HiW * cbar[sl]
    (rows_obj_v,) = red(np.array([rows_obj]))                                  # all-reduce of the objective's row share
    obj = float(rows_obj_v[0]) + fold_obj
    if kind == "kc":
        E = np.stack([(W[:, sl] * cbar[sl]) @ W[:, sl].T for sl in folds])     # stage C
        hb = np.stack([-(W[:, sl] @ mbar[sl]) for sl in folds])
        E, hb = red(E, hb)                                                     # all-reduce C
        for f, sl in enumerate(folds):                                         # stage D
            Hi, h = Hinv[f], hv[f]
            gbar = Hi @ hb[f]
            Y = Hi @ E[f]
            Hbar = -Y @ Hi - 0.5 * (np.outer(gbar, h) + np.outer(h, gbar))
            Wf, lf, af = W[:, sl], lam[sl], alpha[sl]
            abar[sl] = -mbar[sl] * lf + Wf.T @ gbar
            lam_bar[sl] = -mbar[sl] * af + cbar[sl] + np.sum(Wf * (Hbar @ Wf), axis=0) / lf ** 2
            Dm[:, sl] += np.outer(gbar, af) - 2.0 * (Hbar @ Wf) / lf
            GW += np.outer(h, hb[f]) + 2.0 * Y + np.outer(gbar, gv[f]) - 2.0 * Hbar @ P[f]
            beta_bar += -hb[f] - P[f] @ gbar
    tbar = -abar / lam
    lam_bar = lam_bar - abar * alpha / lam
    SW = np.outer(beta, beta_bar) + np.outer(beta_bar, beta) + GW
    C_bar = _phi_adj(LC, -np.tril(solve_triangular(LC, SW, lower=True, trans="T")))
    vy_bar = solve_triangular(LC, beta_bar, lower=True, trans="T")
    CV = C_bar @ V
    lam_bar = lam_bar - (beta_bar @ W) * y / lam ** 2 - np.sum(V * CV, axis=0) / lam ** 2
    Vbar = solve_triangular(LC, np.outer(beta, tbar) + Dm, lower=True, trans="T") \
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
        # block objectives: folds are quarters of the GLOBAL row order; uneven blocks so that a fold straddles the ranks
        gb_ = load_golden("c4_kin_fitc_P1")             # N = 500, fold size 125
        nb_ = gb_["X"].shape[0]
        cuts = [0] + [D.row_block(nb_, r, size)[1] + 37 for r in range(size - 1)] + [nb_]
        blo, bhi = cuts[rank], cuts[rank + 1]

        def ar(t):
            dist.all_reduce(t)

        for kind in ("dss", "kc"):
            val, grad, gU = staged_block(gb_["X"][blo:bhi], gb_["y"][blo:bhi], gb_["U"], gb_["theta"], kind, nb_, blo, ar)
            res[kind] = (abs(val - gb_["obj_" + kind]) / abs(gb_["obj_" + kind]),
                         relerr(grad, grad_vector(gb_, kind)), relerr(gU, gb_["grad_u_" + kind]))
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
        for kind in ("dss", "kc"):               # the sharded block-objective protocol vs the reference's autograd
            assert r[kind][0] <= 1e-8 and r[kind][1] <= 1e-6 and r[kind][2] <= 1e-6, (rank, kind, r[kind])
        assert r["metrics"] <= 1e-7
        assert r["grid"] <= 1e-14
        assert r["grid_shape"] == (4, 5)
