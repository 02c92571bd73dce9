"""Seeded synthetic stand-ins for the reference's data (numpy only).

The KIN40K workbook the scripts read (KF:141, K20:139) is not in the reference
repository, so every config runs on synthetic data of the same shape
(SURVEY.md §8d): 8-D inputs, a smooth target plus 0.1 noise, standardised.
"""
import numpy as np


def kin40k_like(n_train, n_test=0, d=8, seed=0):
    """X ~ N(0,1)^{n x d}, y = sin(Xw) + 0.5 cos(Xv) + 0.1 eps, standardised.

    Train rows come from default_rng(seed), test rows from default_rng(seed+1)
    (same w, v), mirroring the train/test sheets of the workbook (KF:197-200).
    """
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n_train, d))
    w = rng.standard_normal(d) / np.sqrt(d)
    v = rng.standard_normal(d) / np.sqrt(d)
    y = np.sin(X @ w) + 0.5 * np.cos(X @ v) + 0.1 * rng.standard_normal(n_train)
    mu, sd = y.mean(), y.std()
    y = ((y - mu) / sd).reshape(n_train, 1)
    if n_test:
        rng_t = np.random.default_rng(seed + 1)
        Xs = rng_t.standard_normal((n_test, d))
        ys = np.sin(Xs @ w) + 0.5 * np.cos(Xs @ v) + 0.1 * rng_t.standard_normal(n_test)
        ys = ((ys - mu) / sd).reshape(n_test, 1)
        return X, y, Xs, ys
    return X, y


def hyper_point(name, d=8, seed=2):
    """theta = [a, b_1..b_d, c] with a = log sf^2, b = log l, c = log sn^2 (KF:7-12, KF:239).

    P1: script-style initialisation (K20:211-213: para_l ~ U(0,1), para_k = para_noise = 1).
    P2: a "late optimisation" point with small noise, where cond(K) is large.
    """
    if name == "P1":
        b = np.random.default_rng(seed).random(d)
        return np.concatenate([[1.0], b, [1.0]])
    if name == "P2":
        return np.concatenate([[0.0], np.full(d, np.log(2.0)), [np.log(0.01)]])
    raise ValueError(name)


def inducing_init(m, d=8, seed=3):
    """U ~ U(0,1)^{m x d} (K20:215)."""
    return np.random.default_rng(seed).random((m, d))
