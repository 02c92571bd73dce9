"""Size-independent properties of the checker itself (CPU): the oracle restatements must obey the symmetries of the
model they restate — they are what the GPU parity tests trust at sizes and shapes the reference goldens do not cover."""
import math

import numpy as np
import pytest

from oracle import gp_oracle as O
from oracle import woodbury as WB


def _problem(n, d, m, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (n, d))
    y = (np.sin(X @ rng.standard_normal(d)) + 0.1 * rng.standard_normal(n)).reshape(-1, 1)
    U = rng.uniform(-1, 1, (m, d))
    theta = np.concatenate([[0.2], np.log(rng.uniform(0.8, 2.0, d)), [-2.0]])
    return X, y, U, theta


@pytest.mark.parametrize("score", ["crps", "logs", "nlml"])
def test_full_objective_is_invariant_under_row_permutations(score):
    X, y, _, theta = _problem(120, 4, 1, 1)
    p = np.random.default_rng(2).permutation(120)
    v0, g0 = O.full_obj_grad(X, y, theta, O.SCORES[score])
    v1, g1 = O.full_obj_grad(X[p], y[p], theta, O.SCORES[score])
    assert abs(v0 - v1) <= 1e-10 * abs(v0)
    assert np.max(np.abs(g0 - g1)) <= 1e-8 * np.max(np.abs(g0))


@pytest.mark.parametrize("score", ["crps", "logs", "nlml"])
def test_fitc_objective_is_equivariant_under_inducing_point_permutations(score):
    X, y, U, theta = _problem(150, 3, 12, 3)
    q = np.random.default_rng(4).permutation(12)
    v0, g0, gu0 = WB.fitc_obj_grad(X, y, U, theta, O.SCORES[score])[:3]
    v1, g1, gu1 = WB.fitc_obj_grad(X, y, U[q], theta, O.SCORES[score])[:3]
    assert abs(v0 - v1) <= 1e-9 * abs(v0)
    assert np.max(np.abs(g0 - g1)) <= 1e-7 * np.max(np.abs(g0))
    assert np.max(np.abs(gu0[q] - gu1)) <= 1e-7 * np.max(np.abs(gu0))


@pytest.mark.parametrize("kind", ["dss", "kc"])
def test_fitc_block_objectives_depend_on_the_folds_only_as_sets(kind):
    """Rows permuted INSIDE each of the four folds leave the block objectives unchanged; moving rows between folds
    does not (the folds are contiguous quarters of the row order, K20:541-543)."""
    X, y, U, theta = _problem(160, 3, 9, 5)
    rng = np.random.default_rng(6)
    p = np.concatenate([f * 40 + rng.permutation(40) for f in range(4)])
    v0, g0, gu0 = WB.fitc_block_obj_grad(X, y, U, theta, kind)
    v1, g1, gu1 = WB.fitc_block_obj_grad(X[p], y[p], U, theta, kind)
    assert abs(v0 - v1) <= 1e-9 * abs(v0)
    assert np.max(np.abs(g0 - g1)) <= 1e-7 * np.max(np.abs(g0))
    assert np.max(np.abs(gu0 - gu1)) <= 1e-7 * np.max(np.abs(gu0))
    r = rng.permutation(160)
    v2 = WB.fitc_block_obj_grad(X[r], y[r], U, theta, kind)[0]
    assert abs(v2 - v0) > 1e-6 * abs(v0)


def test_nlml_scaling_identity():
    """y -> s y with sf^2, sn^2 -> s^2 sf^2, s^2 sn^2 shifts the NLML by N log s (full GP and FITC)."""
    X, y, U, theta = _problem(90, 2, 8, 7)
    s = 1.7
    th2 = theta.copy()
    th2[0] += 2 * math.log(s)
    th2[-1] += 2 * math.log(s)
    v0 = O.full_obj_grad(X, y, theta, O.SCORE_NLML)[0]
    v1 = O.full_obj_grad(X, s * y, th2, O.SCORE_NLML)[0]
    assert abs(v1 - v0 - 90 * math.log(s)) <= 1e-9 * abs(v0)
    # FITC: the jitter on K_uu does not scale with s, so the identity holds up to O(jitter)
    f0 = WB.fitc_obj_grad(X, y, U, theta, O.SCORE_NLML, jitter=0.0 + 1e-12)[0]
    f1 = WB.fitc_obj_grad(X, s * y, U, th2, O.SCORE_NLML, jitter=(0.0 + 1e-12) * s * s)[0]
    assert abs(f1 - f0 - 90 * math.log(s)) <= 1e-6 * abs(f0)


def test_fitc_with_all_training_points_as_inducing_points_is_the_full_gp():
    """U = X, jitter -> 0: Q_ff = K_ff, the FITC correction vanishes and the LOO scores of the two models agree."""
    X, y, _, theta = _problem(40, 2, 1, 9)
    for score in ("crps", "logs"):
        vf = O.full_obj_grad(X, y, theta, O.SCORES[score])[0]
        vs = WB.fitc_obj_grad(X, y, X.copy(), theta, O.SCORES[score], jitter=1e-10)[0]
        assert abs(vf - vs) <= 1e-5 * abs(vf), score
