"""TEST / BENCH INFRASTRUCTURE — never imported by the product path.

"Reference as written": a float64 torch-autograd restatement of the reference's own CPU path,
operation for operation, so that bench.py can time WHAT THE SCRIPTS DO (not the closed-form
port in gp_oracle.py) on the GPU box's host cores, where /root/reference does not exist.

What the scripts do per optimiser step (kin40k-FULL-compare.py:239-252, the north-star path):

  KF:241  k_ff  = ARD(train_x, train_x, para_k, para_l)          KF:7-23: mm + two bmm + expand + exp
  KF:242  big_k = k_ff + exp(para_noise) * eye(N)
  KF:243  diag( chol_solve(eye(N), big_k) )                      KF:25-29: potrf (upper), then TWO
  KF:244  chol_solve(train_y, big_k)                             general LU solves (gesv) on the
          -> a second potrf + two more gesv                      triangular factors
  KF:245  crps(mean, var, y)                                      KF:60-68
  KF:252  CRPS_ave.backward()                                     autograd through all of the above

torch.potrf / torch.gesv no longer exist; their successors are used exactly as oracle/ref_shim.py
maps them for the golden generation (upper Cholesky factor; LU solve with pivoting).  Pinned by
tests/test_oracle.py::test_ref_as_written_matches_goldens against tests/golden/c3_*.npz, which were
produced by exec-ing the reference's own source text.
"""
import math
import time

import numpy as np
import torch


def _ard(x, xp, a, b):
    """KF:7-23, same operation sequence (scaled inputs, 2 x.xp' - |x|^2 - |xp|^2, exp)."""
    ell = torch.exp(b.view(1, -1))
    xs, xps = x / ell, xp / ell
    n, d = xs.shape
    m = xps.shape[0]
    cross = 2 * torch.mm(xs, xps.transpose(0, 1))
    sq_x = torch.bmm(xs.view(n, 1, d), xs.view(n, d, 1)).view(n, 1).expand(n, m)
    sq_xp = torch.bmm(xps.view(m, 1, d), xps.view(m, d, 1)).view(1, m).expand(n, m)
    return torch.exp(a).expand(n, m) * torch.exp(0.5 * (cross - sq_x - sq_xp))


def _chol_solve(B, A):
    """KF:25-29: upper factor, then two *general* (LU, pivoting) solves on the triangles."""
    c = torch.linalg.cholesky(A).mT          # torch.potrf(A) returned the upper factor
    s1 = torch.linalg.solve(c.transpose(0, 1), B)   # torch.gesv(B, c') [0]
    return torch.linalg.solve(c, s1)                 # torch.gesv(s1, c) [0]


def _crps(m, c, y):
    """KF:60-68."""
    s = c.pow(0.5)
    z = (y - m) / s
    cdf = 0.5 * (1 + torch.erf(z / math.sqrt(2)))
    pdf = (1 / math.sqrt(2 * math.pi)) * torch.exp(-z.pow(2) / 2)
    return (s * (z * (2 * cdf - 1) + 2 * pdf - 1 / math.sqrt(math.pi))).mean()


def _logs(m, c, y):
    """KF:52-57."""
    first = (y - m).pow(2) / (2 * c)
    return (first + c.pow(0.5).log() + 0.5 * np.log(2 * math.pi)).mean()


def full_step(X, y, theta, score="crps"):
    """One loop body KF:239-252 (crps) / KF:416-428 (logs) / KF:329-339 (nlml): forward + backward.
    X [N, D], y [N, 1] numpy; theta = [a, b_1..b_D, c].  Returns (objective, gradient[D+2])."""
    X = torch.as_tensor(np.asarray(X), dtype=torch.float64)
    y = torch.as_tensor(np.asarray(y), dtype=torch.float64).reshape(-1, 1)
    n, d = X.shape
    th = np.asarray(theta, dtype=np.float64)
    para_k = torch.tensor([th[0]], dtype=torch.float64, requires_grad=True)
    para_l = torch.tensor(th[1:-1].reshape(1, -1), dtype=torch.float64, requires_grad=True)
    para_noise = torch.tensor([th[-1]], dtype=torch.float64, requires_grad=True)
    sn2 = torch.exp(para_noise)
    big_k = _ard(X, X, para_k, para_l) + sn2 * torch.eye(n, dtype=torch.float64)
    if score == "nlml":
        # KF:331-334: potrf for the log-determinant, then chol_solve (a second potrf) for the quadratic form
        u = torch.linalg.cholesky(big_k).mT
        obj = (0.5 * n * np.log(2 * math.pi) + torch.log(torch.diag(u)).sum()
               + 0.5 * y.t().mm(_chol_solve(y, big_k))).sum()
    else:
        kinv_diag = torch.diag(_chol_solve(torch.eye(n, dtype=torch.float64), big_k)).view(n, 1)
        mean = y - _chol_solve(y, big_k) / kinv_diag
        var = 1 / kinv_diag
        obj = _crps(mean, var, y) if score == "crps" else _logs(mean, var, y)
    obj.backward()
    g = np.concatenate([para_k.grad.numpy().ravel(), para_l.grad.numpy().ravel(), para_noise.grad.numpy().ravel()])
    return float(obj.detach()), g


def time_full_step(X, y, theta, score="crps", reps=1):
    """Best wall-clock seconds of `reps` forward+backward steps (all torch CPU threads)."""
    best, out = 1e30, None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = full_step(X, y, theta, score)
        best = min(best, time.perf_counter() - t0)
    return best, out
