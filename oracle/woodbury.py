"""TEST INFRASTRUCTURE — algorithm-matched CPU baseline for the FITC path.

O(N M^2) Woodbury restatement of the dense FITC objectives (K20:222-236,
K20:329-344, K20:434-452) with the three-pass analytic adjoint of SURVEY.md
App. A.2, staged exactly like the CUDA path (row pass -> reduce -> M x M
algebra, three times) so that it (a) prototypes the kernels' arithmetic and
(b) gives bench.py an honest like-for-like CPU number: the reference's own
FITC is dense O(N^3), so timing only that would credit the
GPU with the reference's algorithmic waste.

`row_slices` splits the rows the way ranks do; the per-slice accumulators are
summed where the NCCL all-reduce sits in the product.  tests/test_oracle.py
checks this module against the reference-generated goldens.
"""
import math

import numpy as np
from scipy.linalg import cholesky, solve_triangular

from . import gp_oracle as O


def _kern(U, X, a, ell):
    """ARD-SE by direct differences (KF:7-23 up to rounding)."""
    diff = (U[:, None, :] - X[None, :, :]) / ell[None, None, :]
    return math.exp(a) * np.exp(-0.5 * np.sum(diff * diff, axis=2))


def _phi_adj(L, Lbar):
    """Adjoint of a Cholesky factorisation: Abar from Lbar (A = L L')."""
    P = np.tril(L.T @ Lbar)
    P[np.diag_indices_from(P)] *= 0.5
    T = solve_triangular(L, solve_triangular(L, (P + P.T).T, lower=True, trans="T").T, lower=True, trans="T")
    return 0.5 * T


def fitc_obj_grad(X, y, U, theta, score, jitter=O.JITTER, row_slices=None):
    a, b, c = O._split(theta)
    n, D = X.shape
    m = U.shape[0]
    ell = np.exp(np.asarray(b, dtype=np.float64).ravel())
    if ell.size == 1:
        ell = np.full(D, ell[0])
    ea, sn2 = math.exp(a), math.exp(c)
    y = y.reshape(-1)
    if row_slices is None:
        row_slices = [slice(0, n)]

    # replicated: A = K_uu + jitter I = L_A L_A'
    Kuu = _kern(U, U, a, ell)
    LA = cholesky(Kuu + jitter * np.eye(m), lower=True)

    # ---- pass 1 (rows): V, lambda; reduce C, v_y ------------------------------------------
    st = []
    C = np.eye(m)
    vy = np.zeros(m)
    for sl in row_slices:
        Kuf = _kern(U, X[sl], a, ell)                       # M x n_r
        V = solve_triangular(LA, Kuf, lower=True)
        lam = ea - np.sum(V * V, axis=0) + sn2
        C += (V / lam) @ V.T
        vy += V @ (y[sl] / lam)
        st.append(dict(Kuf=Kuf, V=V, lam=lam, y=y[sl], X=X[sl]))
    LC = cholesky(C, lower=True)
    beta = solve_triangular(LC, vy, lower=True)

    # ---- pass 2 (rows): W, d, alpha, score + seeds; reduce beta_bar, R ---------------------
    obj = 0.0
    beta_bar = np.zeros(m)
    R = np.zeros((m, m))
    for s in st:
        W = solve_triangular(LC, s["V"], lower=True)
        r = np.sum(W * W, axis=0)
        lam = s["lam"]
        d = 1.0 / lam - r / lam ** 2
        alpha = (s["y"] - W.T @ beta) / lam
        if score == O.SCORE_NLML:
            obj += 0.5 * np.sum(np.log(lam)) + 0.5 * float(s["y"] @ alpha)
            abar, dbar = 0.5 * s["y"], np.zeros_like(d)
            lam_bar = 0.5 / lam
        else:
            v, abar, dbar = O._score_and_seeds(alpha.reshape(-1, 1) , d.reshape(-1, 1), score)
            # _score_and_seeds normalises by its own length; renormalise to the global N
            k = alpha.shape[0] / n
            obj += v * k
            abar, dbar = abar.ravel() * k, dbar.ravel() * k
            lam_bar = np.zeros_like(lam)
        lam_bar = lam_bar + dbar * (-1.0 / lam ** 2 + 2.0 * r / lam ** 3) - abar * alpha / lam
        rbar = -dbar / lam ** 2
        tbar = -abar / lam
        beta_bar += W @ tbar
        R += (W * rbar) @ W.T
        s.update(W=W, r=r, d=d, alpha=alpha, lam_bar=lam_bar, rbar=rbar, tbar=tbar)
    if score == O.SCORE_NLML:
        obj += 0.5 * n * math.log(2 * math.pi) + np.sum(np.log(np.diag(LC)))
        LC_bar0 = np.diag(1.0 / np.diag(LC))
    else:
        LC_bar0 = np.zeros((m, m))
    SW = np.outer(beta, beta_bar) + 2.0 * R + np.outer(beta_bar, beta)
    LC_bar = -np.tril(solve_triangular(LC, SW, lower=True, trans="T")) + LC_bar0
    C_bar = _phi_adj(LC, LC_bar)
    vy_bar = solve_triangular(LC, beta_bar, lower=True, trans="T")

    # ---- pass 3 (rows): lambda_bar, V_bar, Kuf_bar; reduce S, sums, kernel-gradient partials
    S = np.zeros((m, m))
    sum_lam_bar = 0.0
    g_a = 0.0
    g_b = np.zeros(D)
    g_U = np.zeros((m, D))
    for s in st:
        V, W, lam, ys = s["V"], s["W"], s["lam"], s["y"]
        CV = C_bar @ V
        lam_bar = s["lam_bar"] - (beta_bar @ W) * ys / lam ** 2 - np.sum(V * CV, axis=0) / lam ** 2
        Wbar = np.outer(beta, s["tbar"]) + 2.0 * W * s["rbar"]
        Vbar = solve_triangular(LC, Wbar, lower=True, trans="T") + np.outer(vy_bar, ys / lam) \
            + 2.0 * CV / lam - 2.0 * V * lam_bar
        Kuf_bar = solve_triangular(LA, Vbar, lower=True, trans="T")
        S += Vbar @ V.T
        sum_lam_bar += lam_bar.sum()
        G = Kuf_bar * s["Kuf"]
        g_a += G.sum()
        for dd in range(D):
            diff = U[:, dd][:, None] - s["X"][:, dd][None, :]
            g_b[dd] += np.sum(G * diff * diff) / ell[dd] ** 2
            g_U[:, dd] += -np.sum(G * diff, axis=1) / ell[dd] ** 2
    LA_bar = -np.tril(solve_triangular(LA, S, lower=True, trans="T"))
    A_bar = _phi_adj(LA, LA_bar)
    G2 = A_bar * Kuu
    g_a += ea * sum_lam_bar + G2.sum()
    g_c = sn2 * sum_lam_bar
    for dd in range(D):
        diff = U[:, dd][:, None] - U[:, dd][None, :]
        g_b[dd] += np.sum(G2 * diff * diff) / ell[dd] ** 2
        g_U[:, dd] += -2.0 * np.sum(G2 * diff, axis=1) / ell[dd] ** 2
    loo_mean = np.concatenate([s["y"] - s["alpha"] / s["d"] for s in st])
    loo_var = np.concatenate([1.0 / s["d"] for s in st])
    return float(obj), np.concatenate([[g_a], g_b, [g_c]]), g_U, loo_mean, loo_var


def fitc_predict(X, y, U, Xs, theta, jitter=O.JITTER):
    """K20:76-83 in Woodbury variables: m* = K*u L_A^-T L_C^-T beta,
    var* = sn2 + e^a - |L_A^-1 k_u*|^2 + |L_C^-1 L_A^-1 k_u*|^2."""
    a, b, c = O._split(theta)
    D = X.shape[1]
    m = U.shape[0]
    ell = np.exp(np.asarray(b, dtype=np.float64).ravel())
    if ell.size == 1:
        ell = np.full(D, ell[0])
    ea, sn2 = math.exp(a), math.exp(c)
    LA = cholesky(_kern(U, U, a, ell) + jitter * np.eye(m), lower=True)
    V = solve_triangular(LA, _kern(U, X, a, ell), lower=True)
    lam = ea - np.sum(V * V, axis=0) + sn2
    LC = cholesky(np.eye(m) + (V / lam) @ V.T, lower=True)
    beta = solve_triangular(LC, V @ (y.reshape(-1) / lam), lower=True)
    Vs = solve_triangular(LA, _kern(U, Xs, a, ell), lower=True)
    Ws = solve_triangular(LC, Vs, lower=True)
    mean = Ws.T @ beta
    var = sn2 + ea - np.sum(Vs * Vs, axis=0) + np.sum(Ws * Ws, axis=0)
    return mean.reshape(-1, 1), var.reshape(-1, 1)


def full_obj_grad(X, y, theta, score):
    """Closed-form full GP — identical algorithm to the oracle's dense path;
    re-exported so the CPU baseline has one entry point per model."""
    return O.full_obj_grad(X, y, theta, score)


def fitc_block_obj_grad(X, y, U, theta, kind, jitter=O.JITTER, folds=4):
    """4-fold block-LOO objectives of the FITC model in O(N M^2): "dss" (K20:538-582) and "kc" = block
    CRPS (K20:669-714).  With B = big_Q^-1 = Lam^-1 - Lam^-1 W' W Lam^-1 the fold block is
    B_ff = Lam_f^-1 - (W_f Lam_f^-1)'(W_f Lam_f^-1), so by Woodbury again
        B_ff^-1 = Lam_f + W_f' H_f^-1 W_f,   H_f = I - W_f Lam_f^-1 W_f'   (M x M per fold),
        log|B_ff| = -sum log lam_i + log|H_f|,
    and the fold predictive of row i is  m_i = y_i - lam_i alpha_i - W_i' h_f,  c_i = lam_i + W_i' H_f^-1 W_i
    with g_f = sum_{i in f} W_i alpha_i, h_f = H_f^-1 g_f.  The adjoint re-enters the three-pass chain
    of fitc_obj_grad through per-row seeds (alpha_bar_i, lam_bar_i, D_i = dL/dW_i direct):
        S_W = beta beta_bar' + beta_bar beta' + sum_i D_i W_i',   W_bar_i = t_bar_i beta + D_i."""
    a, b, c = O._split(theta)
    n, D = X.shape
    m = U.shape[0]
    if n % folds:
        raise ValueError("the reference's fold code needs %d | N" % folds)
    nf = n // folds
    ell = np.exp(np.asarray(b, dtype=np.float64).ravel())
    if ell.size == 1:
        ell = np.full(D, ell[0])
    ea, sn2 = math.exp(a), math.exp(c)
    y = y.reshape(-1)
    I = np.eye(m)
    Kuu = _kern(U, U, a, ell)
    LA = cholesky(Kuu + jitter * I, lower=True)
    Kuf = _kern(U, X, a, ell)
    V = solve_triangular(LA, Kuf, lower=True)
    lam = ea - np.sum(V * V, axis=0) + sn2
    LC = cholesky(I + (V / lam) @ V.T, lower=True)
    beta = solve_triangular(LC, V @ (y / lam), lower=True)
    W = solve_triangular(LC, V, lower=True)
    alpha = (y - W.T @ beta) / lam
    # ---- per-fold M x M quantities (pass 2 reductions) --------------------------------------------
    obj = 0.0
    abar = np.zeros(n)
    lam_bar = np.zeros(n)
    Dmat = np.zeros((m, n))              # D_i
    GW = np.zeros((m, m))
    beta_bar = np.zeros(m)
    for f in range(folds):
        sl = slice(f * nf, (f + 1) * nf)
        Wf, lf, af, yf = W[:, sl], lam[sl], alpha[sl], y[sl]
        Pf = (Wf / lf) @ Wf.T                     # I - H_f
        H = I - Pf
        g = Wf @ af
        LH = cholesky(H, lower=True)
        h = solve_triangular(LH, solve_triangular(LH, g, lower=True), lower=True, trans="T")
        Hinv = solve_triangular(LH, solve_triangular(LH, I, lower=True), lower=True, trans="T")
        if kind == "dss":
            obj += 0.5 * nf * math.log(2 * math.pi) + 0.5 * np.sum(np.log(lf)) - np.sum(np.log(np.diag(LH))) \
                + 0.5 * np.sum(lf * af * af) + 0.5 * float(g @ h)
            Hhat = -0.5 * Hinv - 0.5 * np.outer(h, h)
            abar[sl] = lf * af + Wf.T @ h
            lam_bar[sl] = 0.5 / lf + 0.5 * af * af + np.sum(Wf * (Hhat @ Wf), axis=0) / lf ** 2
            Dmat[:, sl] = -2.0 * (Hhat @ Wf) / lf + np.outer(h, af)
            GW += -2.0 * Hhat @ Pf + np.outer(h, g)
            beta_bar += -h
        else:
            HiW = Hinv @ Wf                                   # H_f^-1 W_i
            mu = yf - lf * af - Wf.T @ h
            cv = lf + np.sum(Wf * HiW, axis=0)
            sd = np.sqrt(cv)
            z = (yf - mu) / sd
            tpm1 = 2 * O._Phi(z) - 1
            obj += float(np.mean(sd * (z * tpm1 + 2 * O._phi(z) - 1 / math.sqrt(math.pi))))
            mbar = -tpm1 / nf
            cbar = (2 * O._phi(z) - 1 / math.sqrt(math.pi)) / (2 * sd) / nf
            # pass 2b reductions
            hbar = -(Wf @ mbar)
            E = (Wf * cbar) @ Wf.T                            # sum cbar_i W_i W_i'
            gbar = Hinv @ hbar
            Hbar = -Hinv @ E @ Hinv - 0.5 * (np.outer(gbar, h) + np.outer(h, gbar))
            abar[sl] = -mbar * lf + Wf.T @ gbar
            lam_bar[sl] = -mbar * af + cbar + np.sum(Wf * (Hbar @ Wf), axis=0) / lf ** 2
            Dmat[:, sl] = -np.outer(h, mbar) + 2.0 * HiW * cbar + np.outer(gbar, af) - 2.0 * (Hbar @ Wf) / lf
            GW += np.outer(h, hbar) + 2.0 * Hinv @ E + np.outer(gbar, g) - 2.0 * Hbar @ Pf
            beta_bar += -hbar - Pf @ gbar
    # ---- re-enter the three-pass adjoint ------------------------------------------------------------
    tbar = -abar / lam
    lam_bar = lam_bar - abar * alpha / lam
    SW = np.outer(beta, beta_bar) + np.outer(beta_bar, beta) + GW
    LC_bar = -np.tril(solve_triangular(LC, SW, lower=True, trans="T"))
    C_bar = _phi_adj(LC, LC_bar)
    vy_bar = solve_triangular(LC, beta_bar, lower=True, trans="T")
    CV = C_bar @ V
    lam_bar = lam_bar - (beta_bar @ W) * y / lam ** 2 - np.sum(V * CV, axis=0) / lam ** 2
    Wbar = np.outer(beta, tbar) + Dmat
    Vbar = solve_triangular(LC, Wbar, lower=True, trans="T") + np.outer(vy_bar, y / lam) + 2.0 * CV / lam - 2.0 * V * lam_bar
    Kuf_bar = solve_triangular(LA, Vbar, lower=True, trans="T")
    S = Vbar @ V.T
    G = Kuf_bar * Kuf
    LA_bar = -np.tril(solve_triangular(LA, S, lower=True, trans="T"))
    G2 = _phi_adj(LA, LA_bar) * Kuu
    g_a = ea * lam_bar.sum() + G.sum() + G2.sum()
    g_c = sn2 * lam_bar.sum()
    g_b = np.zeros(D)
    g_U = np.zeros((m, D))
    for dd in range(D):
        dux = U[:, dd][:, None] - X[:, dd][None, :]
        duu = U[:, dd][:, None] - U[:, dd][None, :]
        g_b[dd] = (np.sum(G * dux * dux) + np.sum(G2 * duu * duu)) / ell[dd] ** 2
        g_U[:, dd] = (-np.sum(G * dux, axis=1) - 2.0 * np.sum(G2 * duu, axis=1)) / ell[dd] ** 2
    return float(obj), np.concatenate([[g_a], g_b, [g_c]]), g_U


def fitc_obj_grad_kspace(X, y, U, theta, score, jitter=O.JITTER, row_slices=None):
    """Arithmetic prototype of the fused row kernels (csrc/gps_fitc_fused.cu): passes 1 and 2 as in
    `fitc_obj_grad` with explicit triangular inverses (V = L_A^-1 k, W = T2 k, T2 = L_C^-1 L_A^-1), pass 3
    in "k-space": per row only k_i is needed,
        Kuf_bar_i = t_i a1 + (y_i/lam_i) a2 + [2 r_i E1 + (2/lam_i) E2 - 2 lam_bar_i E3] k_i
        E1 = T2'T2,  E2 = L_A^-T C_bar L_A^-1,  E3 = A^-1,  a1 = T2' beta,  a2 = L_A^-T vy_bar,
    the row reductions are Z = sum lam_bar_i k_i k_i', P = sum G_i [x_i | 1], sum_i x_id^2 sum_m G_im, and
    the S = sum V_bar_i V_i' that the Cholesky adjoint of L_A needs is assembled from the M x M accumulators
    of all three passes:
        S = b1 (L_C beta_bar)' + vy_bar vy' + 2 L_C^-T R L_C' + 2 C_bar (C - I) - 2 L_A^-1 Z L_A^-T.
    Same values and gradients as `fitc_obj_grad` up to rounding (tests/test_oracle.py)."""
    a, b, c = O._split(theta)
    n, D = X.shape
    m = U.shape[0]
    ell = np.exp(np.asarray(b, dtype=np.float64).ravel())
    if ell.size == 1:
        ell = np.full(D, ell[0])
    ea, sn2 = math.exp(a), math.exp(c)
    y = y.reshape(-1)
    if row_slices is None:
        row_slices = [slice(0, n)]
    Us = U / ell
    Kuu = _kern(U, U, a, ell)
    LA = cholesky(Kuu + jitter * np.eye(m), lower=True)
    LAi = solve_triangular(LA, np.eye(m), lower=True)
    # pass 1
    Cm = np.zeros((m, m))
    vy = np.zeros(m)
    st = []
    for sl in row_slices:
        K = _kern(U, X[sl], a, ell)
        V = LAi @ K
        lam = ea - np.sum(V * V, axis=0) + sn2
        Cm += (V / lam) @ V.T
        vy += V @ (y[sl] / lam)
        st.append(dict(K=K, lam=lam, y=y[sl], Xs=X[sl] / ell))
    LC = cholesky(np.eye(m) + Cm, lower=True)
    LCi = solve_triangular(LC, np.eye(m), lower=True)
    beta = LCi @ vy
    T2 = LCi @ LAi
    c2 = T2.T @ beta
    # pass 2
    obj = 0.0
    beta_bar = np.zeros(m)
    R = np.zeros((m, m))
    for s in st:
        W = T2 @ s["K"]
        r = np.sum(W * W, axis=0)
        lam = s["lam"]
        d = 1.0 / lam - r / lam ** 2
        alpha = (s["y"] - c2 @ s["K"]) / lam
        if score == O.SCORE_NLML:
            obj += 0.5 * np.sum(np.log(lam)) + 0.5 * float(s["y"] @ alpha)
            abar, dbar = 0.5 * s["y"], np.zeros_like(d)
            lb0 = 0.5 / lam
        else:
            v, abar, dbar = O._score_and_seeds(alpha.reshape(-1, 1), d.reshape(-1, 1), score)
            kk = alpha.shape[0] / n
            obj += v * kk
            abar, dbar = abar.ravel() * kk, dbar.ravel() * kk
            lb0 = np.zeros_like(lam)
        lb0 = lb0 + dbar * (-1.0 / lam ** 2 + 2.0 * r / lam ** 3) - abar * alpha / lam
        rbar = -dbar / lam ** 2
        tbar = -abar / lam
        beta_bar += W @ tbar
        R += (W * rbar) @ W.T
        s.update(lb0=lb0, rbar=rbar, tbar=tbar, alpha=alpha, d=d)
    if score == O.SCORE_NLML:
        obj += 0.5 * n * math.log(2 * math.pi) + np.sum(np.log(np.diag(LC)))
        LC_bar0 = np.diag(1.0 / np.diag(LC))
    else:
        LC_bar0 = np.zeros((m, m))
    SW = np.outer(beta, beta_bar) + 2.0 * R + np.outer(beta_bar, beta)
    LC_bar = -np.tril(LCi.T @ SW) + LC_bar0
    C_bar = _phi_adj(LC, LC_bar)
    vy_bar = LCi.T @ beta_bar
    E1 = T2.T @ T2
    E2 = LAi.T @ C_bar @ LAi
    E3 = LAi.T @ LAi
    a1 = c2
    a2 = LAi.T @ vy_bar
    c1 = T2.T @ beta_bar
    # pass 3 (k-space)
    Z = np.zeros((m, m))
    P = np.zeros((m, D))
    S0 = np.zeros(m)
    xq = np.zeros(D)
    sum_lb = 0.0
    for s in st:
        K, lam, ys = s["K"], s["lam"], s["y"]
        P2 = E2 @ K
        s1 = np.sum(K * P2, axis=0)
        bw = c1 @ K
        lb = s["lb0"] - bw * ys / lam ** 2 - s1 / lam ** 2
        Kb = np.outer(a1, s["tbar"]) + np.outer(a2, ys / lam) + 2.0 * (E1 @ K) * s["rbar"] + 2.0 * P2 / lam \
            - 2.0 * (E3 @ K) * lb
        G = Kb * K
        Z += (K * lb) @ K.T
        P += G @ s["Xs"]
        S0 += G.sum(axis=1)
        xq += (s["Xs"] ** 2).T @ G.sum(axis=0)
        sum_lb += lb.sum()
    # finish
    b1 = LCi.T @ beta
    S = np.outer(b1, LC @ beta_bar) + np.outer(vy_bar, vy) + 2.0 * LCi.T @ R @ LC.T + 2.0 * C_bar @ Cm \
        - 2.0 * LAi @ Z @ LAi.T
    LA_bar = -np.tril(LAi.T @ S)
    A_bar = _phi_adj(LA, LA_bar)
    G2 = A_bar * Kuu
    g_a = ea * sum_lb + G2.sum() + S0.sum()
    g_c = sn2 * sum_lb
    g_b = np.zeros(D)
    g_U = np.zeros((m, D))
    for dd in range(D):
        u = Us[:, dd]
        diff = u[:, None] - u[None, :]
        g_b[dd] = np.sum(u * u * S0) - 2.0 * np.sum(u * P[:, dd]) + xq[dd] + np.sum(G2 * diff * diff)
        g_U[:, dd] = (-(u * S0 - P[:, dd]) - 2.0 * np.sum(G2 * diff, axis=1)) / ell[dd]
    loo_mean = np.concatenate([s["y"] - s["alpha"] / s["d"] for s in st])
    loo_var = np.concatenate([1.0 / s["d"] for s in st])
    return float(obj), np.concatenate([[g_a], g_b, [g_c]]), g_U, loo_mean, loo_var
