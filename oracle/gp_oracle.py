"""TEST INFRASTRUCTURE — CPU oracle for the GP scoring-rule hot path.

A numpy/scipy float64 restatement of the reference's algorithm, function by
function, each citing the reference file:line it follows (abbreviations as in
SURVEY.md: KF = kin40k-FULL-compare.py, K20 = KIN40K-COMPARE-ALL-FITC-20.py,
SF/SC = the SIMPLE scripts, CP = contour-plot.R).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may import this module.  The product path (the CUDA library behind
include/gpscore.h) never does and has no CPU fallback.

Pinning.  The reference has no tests or golden vectors (SURVEY.md §4, §8c:
"parity unpinned" at the reference).  The pin is tests/golden/*.npz: outputs of
the reference's own source text, exec-ed from /root/reference by
tests/golden/make_golden.py in float64.  tests/test_oracle.py checks every
function here against those files.

The reference differentiates with torch autograd (KF:252).  The oracle uses the
closed-form adjoints of the same *dense* computation:
the FITC functions below factor the dense N x N `big_Q` exactly as K20:222-231
does — they do NOT use the Woodbury identity the CUDA path uses, so oracle and
product are independent derivations of the same numbers.
"""
import math

import numpy as np
from scipy.linalg import cho_factor, cho_solve
from scipy.special import erf

SCORE_CRPS, SCORE_LOGS, SCORE_NLML, SCORE_DSS = 0, 1, 2, 3
SCORES = {"crps": SCORE_CRPS, "logs": SCORE_LOGS, "nlml": SCORE_NLML, "dss": SCORE_DSS}
JITTER = 1e-3  # K20:36, part of the model


# ----------------------------------------------------------------------------------------------
# L1: kernel and linear-algebra helpers
# ----------------------------------------------------------------------------------------------
def ARD(x, xp, a, b):
    """KF:7-23. e^a * exp(-0.5*||(x - x')/e^b||^2) through the same norm expansion
    (2 x.x' - |x|^2 - |x'|^2); b is log l (KF:10-12), broadcast if it has one element (KF:8)."""
    b = np.asarray(b, dtype=np.float64).reshape(1, -1)
    ell = np.exp(b)
    xs = x / ell
    xps = xp / ell
    res = 2.0 * xs @ xps.T
    res = res - np.sum(xs * xs, axis=1)[:, None] - np.sum(xps * xps, axis=1)[None, :]
    return np.exp(a) * np.exp(0.5 * res)


def chol_solve(B, A):
    """KF:25-29. A^{-1} B through a Cholesky factor of A."""
    return cho_solve(cho_factor(A, lower=True), B)


def Q(a_pts, u, b_pts, para_k, para_l, jitter=JITTER):
    """KF:32-39. K_au (K_uu + 1e-3 I)^{-1} K_ub."""
    K_au = ARD(a_pts, u, para_k, para_l)
    K_uu = ARD(u, u, para_k, para_l) + jitter * np.eye(u.shape[0])
    K_ub = ARD(u, b_pts, para_k, para_l)
    return K_au @ chol_solve(K_ub, K_uu)


# ----------------------------------------------------------------------------------------------
# L2: scoring rules and metrics
# ----------------------------------------------------------------------------------------------
def _Phi(z):
    return 0.5 * (1.0 + erf(z / math.sqrt(2.0)))


def _phi(z):
    return np.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi)


def crps(m, c, y):
    """KF:60-68. Mean closed-form Gaussian CRPS; c is the predictive VARIANCE (sqrt at KF:63)."""
    s = np.sqrt(c)
    z = (y - m) / s
    return float(np.mean(s * (z * (2.0 * _Phi(z) - 1.0) + 2.0 * _phi(z) - 1.0 / math.sqrt(math.pi))))


def logs(m, c, y):
    """KF:52-57. Mean Gaussian negative log predictive density."""
    return float(np.mean((y - m) ** 2 / (2.0 * c) + 0.5 * np.log(c) + 0.5 * math.log(2.0 * math.pi)))


def SMSE(m, y, y_train):
    """KF:128-134."""
    return float(np.mean((m - y) ** 2) / np.mean((y_train.mean() - y) ** 2))


def trivial_loss(m, c, y, y_train):
    """KF:110-119 (MSLL); var() is the unbiased estimator (KF:114)."""
    mu, var = y_train.mean(), y_train.var(ddof=1)
    ls = (y - m) ** 2 / (2.0 * c) + 0.5 * np.log(c) + 0.5 * math.log(2.0 * math.pi)
    triv = 0.5 * math.log(2.0 * math.pi * var) + (y - mu) ** 2 / (2.0 * var)
    return float(np.mean(ls - triv))


def coverage(m, c, y):
    """KF:288-292. Fraction of y strictly inside mean +- 2 sd."""
    s = np.sqrt(c)
    return float(np.mean(((m + 2 * s - y) > 0) & ((y - (m - 2 * s)) > 0)))


def test_metrics(m, c, y, y_train):
    """KF:276-292 in one dict."""
    return dict(mse=float(np.mean((m - y) ** 2)), smse=SMSE(m, y, y_train), logs=logs(m, c, y),
                crps=crps(m, c, y), msll=trivial_loss(m, c, y, y_train), coverage=coverage(m, c, y))


# ----------------------------------------------------------------------------------------------
# L3: objectives (dense, as the reference assembles them) and their closed-form adjoints
# ----------------------------------------------------------------------------------------------
def _split(theta):
    theta = np.asarray(theta, dtype=np.float64).ravel()
    return theta[0], theta[1:-1], theta[-1]


def loo_transform(K, y):
    """KF:241-244. d = diag(K^-1), alpha = K^-1 y, mu = y - alpha/d, s2 = 1/d."""
    B = chol_solve(np.eye(K.shape[0]), K)
    d = np.diag(B).reshape(-1, 1)
    alpha = B @ y
    return y - alpha / d, 1.0 / d, B, alpha, d


def _score_and_seeds(alpha, d, score):
    """Value of the LOO score as a function of (alpha, d) and dL/dalpha, dL/dd."""
    n = alpha.shape[0]
    if score == SCORE_CRPS:
        z = alpha / np.sqrt(d)
        g = z * (2 * _Phi(z) - 1) + 2 * _phi(z) - 1 / math.sqrt(math.pi)
        val = np.mean(g / np.sqrt(d))
        abar = (2 * _Phi(z) - 1) / (n * d)
        dbar = -(0.5 * d ** -1.5 * g + 0.5 * (2 * _Phi(z) - 1) * alpha * d ** -2.0) / n
    elif score == SCORE_LOGS:
        val = np.mean(alpha ** 2 / (2 * d) - 0.5 * np.log(d) + 0.5 * math.log(2 * math.pi))
        abar = alpha / (n * d)
        dbar = -(alpha ** 2 / (2 * d ** 2) + 1 / (2 * d)) / n
    else:
        raise ValueError(score)
    return float(val), abar, dbar


def _dL_dK(K, y, score):
    """Objective value and W = dL/dK (symmetric) for K including the noise term."""
    n = K.shape[0]
    if score == SCORE_NLML:
        # KF:331-334: 0.5 N log 2pi + sum log diag(chol) + 0.5 y' K^-1 y
        c, low = cho_factor(K, lower=True)
        alpha = cho_solve((c, low), y)
        val = 0.5 * n * math.log(2 * math.pi) + np.sum(np.log(np.diag(c))) + 0.5 * float((y.T @ alpha).item())
        B = cho_solve((c, low), np.eye(n))
        return float(val), 0.5 * (B - alpha @ alpha.T), None, None
    mu, s2, B, alpha, d = loo_transform(K, y)
    val, abar, dbar = _score_and_seeds(alpha, d, score)
    u = B @ abar
    W = -(0.5 * (u @ alpha.T + alpha @ u.T) + (B * dbar.T) @ B)
    return val, W, mu, s2


def full_objective(X, y, theta, score):
    """KF:239-245 (crps), KF:416-424 (logs), KF:329-334 (nlml): value, LOO mean, LOO variance."""
    a, b, c = _split(theta)
    K = ARD(X, X, a, b) + math.exp(c) * np.eye(X.shape[0])
    val, _, mu, s2 = _dL_dK(K, y, score)
    return val, mu, s2


def _kernel_param_grads(Wk, Kmat, x, xp, b, same):
    """sum_ij Wk_ij dK_ij/d{a, b_d} for K = ARD(x, xp)."""
    G = Wk * Kmat
    ga = G.sum()
    ell2 = np.exp(2 * np.asarray(b, dtype=np.float64).ravel())
    D = x.shape[1]
    if ell2.size == 1:
        ell2 = np.full(D, ell2[0])
    gb = np.empty(D)
    for dd in range(D):
        diff2 = (x[:, dd][:, None] - xp[:, dd][None, :]) ** 2
        gb[dd] = np.sum(G * diff2) / ell2[dd]
    return ga, gb, G


def full_obj_grad(X, y, theta, score):
    """Value and gradient wrt theta = [a, b_1..b_D, c] of the full-GP objectives
    (what KF:252 / KF:339 / KF:428 obtain from autograd)."""
    a, b, c = _split(theta)
    Kf = ARD(X, X, a, b)
    K = Kf + math.exp(c) * np.eye(X.shape[0])
    val, W, _, _ = _dL_dK(K, y, score)
    ga, gb, _ = _kernel_param_grads(W, Kf, X, X, b, True)
    gc = math.exp(c) * np.trace(W)
    return val, np.concatenate([[ga], gb, [gc]])


def dss(m, c, shape1, y):
    """KF:103-108. Multivariate Gaussian negative log density (Dawid-Sebastiani form) of one fold."""
    cf, low = cho_factor(c, lower=True)
    r = y - m
    return float(0.5 * shape1 * math.log(2 * math.pi) + np.sum(np.log(np.diag(cf)))
                 + 0.5 * (r.T @ cho_solve((cf, low), r)).item())


def _block_dL_dK(K, y, kind, folds=4):
    """Value and W = dL/dK of the 4-fold block-LOO objectives on a matrix K that already includes the
    noise: kind "dss" (KF:499-538, K20:538-582) or "kc" = block CRPS (K20:669-714).  With B = K^-1,
    C_f = B_ff^-1: m_f = y_f - C_f (B y)_f, cov_f = C_f (dss) or diag(C_f) (kc)."""
    n = K.shape[0]
    if n % folds:
        raise ValueError("the reference's fold code needs %d | N" % folds)
    B = chol_solve(np.eye(n), K)
    alpha = B @ y
    nf = n // folds
    val = 0.0
    Gamma = np.zeros((n, n))
    abar = np.zeros((n, 1))
    for f in range(folds):
        sl = slice(f * nf, (f + 1) * nf)
        cf, low = cho_factor(B[sl, sl], lower=True)
        Cf = cho_solve((cf, low), np.eye(nf))
        if kind == "dss":
            ab = Cf @ alpha[sl]
            val += 0.5 * nf * math.log(2 * math.pi) - np.sum(np.log(np.diag(cf))) + 0.5 * (alpha[sl].T @ ab).item()
            Gamma[sl, sl] = -0.5 * (Cf + ab @ ab.T)
            abar[sl] = ab
        else:
            m = y[sl] - Cf @ alpha[sl]
            c = np.diag(Cf).reshape(-1, 1)
            sd = np.sqrt(c)
            z = (y[sl] - m) / sd
            val += float(np.mean(sd * (z * (2 * _Phi(z) - 1) + 2 * _phi(z) - 1 / math.sqrt(math.pi))))
            mbar = -(2 * _Phi(z) - 1) / nf
            cbar = (2 * _phi(z) - 1 / math.sqrt(math.pi)) / (2 * sd) / nf
            Cbar = np.diag(cbar.ravel()) - 0.5 * (mbar @ alpha[sl].T + alpha[sl] @ mbar.T)
            Gamma[sl, sl] = -Cf @ Cbar @ Cf
            abar[sl] = -Cf @ mbar
    u = B @ abar
    W = -(B @ Gamma @ B + 0.5 * (u @ alpha.T + alpha @ u.T))
    return float(val), W


def full_block_obj_grad(X, y, theta, kind):
    """4-fold DSS / kc objective and gradient for the full GP."""
    a, b, c = _split(theta)
    Kf = ARD(X, X, a, b)
    val, W = _block_dL_dK(Kf + math.exp(c) * np.eye(X.shape[0]), y, kind)
    ga, gb, _ = _kernel_param_grads(W, Kf, X, X, b, True)
    return val, np.concatenate([[ga], gb, [math.exp(c) * np.trace(W)]])


def full_dss_obj_grad(X, y, theta, folds=4):
    """KF:499-538: 4-fold block-LOO DSS.  With B = K^-1 the fold predictive is
    m_f = y_f - B_ff^-1 (B y)_f, cov_f = B_ff^-1 (KF:508-530), so
    dss_f = n_f/2 log 2pi - 1/2 log|B_ff| + 1/2 a_f' B_ff^-1 a_f with a = B y.
    Gradient by the adjoint of the dense computation: dL/dB_ff = -1/2 (C_f + abar_f abar_f'),
    C_f = B_ff^-1, abar_f = C_f a_f; dL/dK = -(B Gamma B + sym(u a')), u = B abar.
    The script sizes every fold with index1 = N/4 (KF:521-530), i.e. it assumes 4 | N."""
    a, b, c = _split(theta)
    n = X.shape[0]
    if n % folds:
        raise ValueError("the reference's fold code needs %d | N" % folds)
    Kf = ARD(X, X, a, b)
    K = Kf + math.exp(c) * np.eye(n)
    B = chol_solve(np.eye(n), K)
    alpha = B @ y
    nf = n // folds
    val = 0.0
    Gamma = np.zeros((n, n))
    abar = np.zeros((n, 1))
    for f in range(folds):
        sl = slice(f * nf, (f + 1) * nf)
        Bff = B[sl, sl]
        cf, low = cho_factor(Bff, lower=True)
        Cf = cho_solve((cf, low), np.eye(nf))
        ab = Cf @ alpha[sl]
        val += 0.5 * nf * math.log(2 * math.pi) - np.sum(np.log(np.diag(cf))) + 0.5 * (alpha[sl].T @ ab).item()
        Gamma[sl, sl] = -0.5 * (Cf + ab @ ab.T)
        abar[sl] = ab
    u = B @ abar
    W = -(B @ Gamma @ B + 0.5 * (u @ alpha.T + alpha @ u.T))
    ga, gb, _ = _kernel_param_grads(W, Kf, X, X, b, True)
    gc = math.exp(c) * np.trace(W)
    return float(val), np.concatenate([[ga], gb, [gc]])


def fitc_bigQ(X, U, theta, jitter=JITTER):
    """K20:223-229. Dense N x N big_Q = Q_ff + diag(diag(K_ff - Q_ff + sn2 I))."""
    a, b, c = _split(theta)
    n = X.shape[0]
    k_ff = ARD(X, X, a, b)
    Q_ff = Q(X, U, X, a, b, jitter)
    G = np.diag(np.diag(k_ff - Q_ff + math.exp(c) * np.eye(n)))
    return Q_ff + G, k_ff, Q_ff


def fitc_objective(X, y, U, theta, score, jitter=JITTER):
    """K20:222-234 (crps), K20:434-447 (logs), K20:329-340 (nlml)."""
    bigQ, _, _ = fitc_bigQ(X, U, theta, jitter)
    val, _, mu, s2 = _dL_dK(bigQ, y, score)
    return val, mu, s2


def fitc_obj_grad(X, y, U, theta, score, jitter=JITTER):
    """Value and gradients (theta[D+2], U[M,D]) of the dense FITC objectives
    (what K20:236 / K20:344 / K20:452 obtain from autograd), by the chain rule through
    big_Q = Q_ff + diag(k_ii - Q_ii + sn2),  Q_ff = K_fu A^-1 K_uf,  A = K_uu + jitter I."""
    a, b, c = _split(theta)
    n, D = X.shape
    m = U.shape[0]
    Kuf = ARD(U, X, a, b)
    Kuu = ARD(U, U, a, b)
    A = Kuu + jitter * np.eye(m)
    AiKuf = chol_solve(Kuf, A)
    Q_ff = Kuf.T @ AiKuf
    lam = math.exp(a) - np.diag(Q_ff) + math.exp(c)   # diag(k_ff) = e^a
    bigQ = Q_ff + np.diag(lam)        # K20:225-229
    if score in ("dss", "kc"):
        val, W = _block_dL_dK(bigQ, y, score)   # K20:538-582, K20:669-714
    else:
        val, W, _, _ = _dL_dK(bigQ, y, score)
    Wd = np.diag(W).copy()
    Wq = W - np.diag(Wd)              # dL/dQ_ff: the diagonal of big_Q does not depend on Q_ff
    T = AiKuf @ Wq                    # M x N
    Kuf_bar = 2.0 * T
    A_bar = -T @ AiKuf.T
    ga1, gb1, G1 = _kernel_param_grads(Kuf_bar, Kuf, U, X, b, False)
    ga2, gb2, G2 = _kernel_param_grads(A_bar, Kuu, U, U, b, True)
    ga = ga1 + ga2 + math.exp(a) * Wd.sum()
    gb = gb1 + gb2
    gc = math.exp(c) * Wd.sum()
    ell2 = np.exp(2 * np.asarray(b, dtype=np.float64).ravel())
    if ell2.size == 1:
        ell2 = np.full(D, ell2[0])
    gU = np.empty((m, D))
    for dd in range(D):
        dUX = U[:, dd][:, None] - X[:, dd][None, :]
        dUU = U[:, dd][:, None] - U[:, dd][None, :]
        gU[:, dd] = (-(G1 * dUX).sum(axis=1) - 2.0 * (G2 * dUU).sum(axis=1)) / ell2[dd]
    return val, np.concatenate([[ga], gb, [gc]]), gU, lam


# ----------------------------------------------------------------------------------------------
# L4: prediction
# ----------------------------------------------------------------------------------------------
def cal_mean_and_cov(k1, k2, k3, num, eye_num, data_y, sigma_noise_sq):
    """KF:121-126 (sigma_noise_sq is a module global there)."""
    Kn = k2 + sigma_noise_sq * np.eye(eye_num)
    mean = k1 @ chol_solve(data_y, Kn)
    cov = sigma_noise_sq * np.eye(num) + k3 - k1 @ chol_solve(k1.T, Kn)
    return mean, cov


def spgp_cal_mean_and_cov(k1, Q1, Q2, k2, num_test, num_jitter, data_y, sigma_noise_sq):
    """K20:76-83."""
    G = np.diag(np.diag(k1 - Q1 + sigma_noise_sq * np.eye(num_jitter)))
    mean = Q2 @ chol_solve(data_y, Q1 + G)
    cov = sigma_noise_sq * np.eye(num_test) + k2 - Q2 @ chol_solve(Q2.T, Q1 + G)
    return mean, cov


def full_predict(X, y, Xs, theta):
    """KF:267-273: predictive mean and the DIAGONAL of the covariance."""
    a, b, c = _split(theta)
    mean, cov = cal_mean_and_cov(ARD(Xs, X, a, b), ARD(X, X, a, b), ARD(Xs, Xs, a, b),
                                 Xs.shape[0], X.shape[0], y, math.exp(c))
    return mean, np.diag(cov).reshape(-1, 1)


def fitc_predict(X, y, U, Xs, theta, jitter=JITTER):
    """K20:270-277."""
    a, b, c = _split(theta)
    mean, cov = spgp_cal_mean_and_cov(ARD(X, X, a, b), Q(X, U, X, a, b, jitter), Q(Xs, U, X, a, b, jitter),
                                      ARD(Xs, Xs, a, b), Xs.shape[0], X.shape[0], y, math.exp(c))
    return mean, np.diag(cov).reshape(-1, 1)


# ----------------------------------------------------------------------------------------------
# Grid functions: the R twins contour-plot.R evaluates on the 50 x 50 (l, noise sd) grid
# ----------------------------------------------------------------------------------------------
def rbf_R(x1, x2, l=1.0, k=1.0):
    """CP:15-23. k^2 exp(-0.5 ((x - x')/l)^2), natural parameters, 1-D inputs."""
    x1 = np.asarray(x1, dtype=np.float64).ravel()
    x2 = np.asarray(x2, dtype=np.float64).ravel()
    return k * k * np.exp(-0.5 * ((x1[:, None] - x2[None, :]) / l) ** 2)


def m_CRPS(y, mu, sigma):
    """CP:3-8 — takes the standard deviation, unlike Python crps (KF:63)."""
    return crps(mu, sigma ** 2, y)


def m_Log_score(y, mu, sigma):
    """CP:10-13."""
    return logs(mu, sigma ** 2, y)


def cal_m_crps(x, y, i, j):
    """CP:43-53. LOO-CRPS at length scale i, noise s.d. j."""
    n = len(x)
    big_k = rbf_R(x, x, l=i) + np.eye(n) * j * j
    mu, s2, _, _, _ = loo_transform(big_k, y)
    return m_CRPS(y, mu, np.sqrt(s2))


def wrong_cal_m_crps(x, y, i, j):
    """CP:55-64. In-sample ("wrong") CRPS: scores the posterior at the training points."""
    n = len(x)
    k_ff = rbf_R(x, x, l=i)
    Kn = k_ff + np.eye(n) * j * j
    mean = k_ff @ chol_solve(y, Kn)
    cov = np.diag(np.eye(n) * j * j + k_ff - k_ff @ chol_solve(k_ff, Kn)).reshape(-1, 1)
    return m_CRPS(y, mean, np.sqrt(cov))


def cal_NLML(x, y, i, j):
    """CP:68-73 (R uses log(det(.)) directly; the log-determinant is taken from the Cholesky
    factor here, which equals it whenever det does not underflow)."""
    n = len(x)
    Kn = rbf_R(x, x, l=i) + np.eye(n) * j * j
    c, low = cho_factor(Kn, lower=True)
    return float((0.5 * y.T @ cho_solve((c, low), y)).item() + np.sum(np.log(np.diag(c))) + n / 2 * math.log(2 * math.pi))


def cal_m_logs(x, y, i, j):
    """CP:75-85. LOO log score; adds j^2 to the LOO variance a second time (CP:81)."""
    n = len(x)
    big_k = rbf_R(x, x, l=i) + np.eye(n) * j * j
    mu, s2, _, _, _ = loo_transform(big_k, y)
    return m_Log_score(y, mu, np.sqrt(s2 + j * j))
