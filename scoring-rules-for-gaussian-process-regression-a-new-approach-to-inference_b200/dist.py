"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing (NCCL over
NVLink on the box, gloo in the CPU tests).  Only the sub-paths that shard naturally are sharded
(SURVEY.md §8e):

  FITC objective+gradient   rows of X — three all-reduces of small packed accumulators per evaluation
  prediction + test scoring rows of the test set — one all-reduce of the six metric sums
  grid / restart sweeps     grid points round-robin — results combined by an all-reduce of a
                            zero-filled vector
  full-GP objective         does NOT shard ("replicas only"): one evaluation per GPU

The compute behind each step is an object with the staged FITC protocol of include/gpscore.h
(begin / pass1 / pass2 / pass3 / finish); in production that is `api.Context`.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def row_block(n, rank, size):
    """Contiguous block [lo, hi) of n rows owned by `rank`: sizes differ by at most one."""
    lo = (n * rank) // size
    hi = (n * (rank + 1)) // size
    return lo, hi


def round_robin(n, rank, size):
    """Indices of the grid / restart points owned by `rank`."""
    return np.arange(rank, n, size)


def allreduce_sum_(t, group=None):
    """In-place sum of a 1-D float64 tensor across ranks (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class ShardedFitc:
    """Row-sharded FITC evaluation.  `backend` holds THIS rank's rows and implements
    fitc_eval_sharded(theta, U, score, world_n, allreduce) -> (obj, g_theta, g_U).

    library_comm=True (production, api.Context after `comm_init`): the three all-reduces are issued by
    libgpscore itself on the context's stream (NCCL inside the call, no host synchronisation between the
    passes).  Otherwise the staged protocol runs with torch.distributed doing the all-reduces (gloo in
    the CPU tests, with a numpy stand-in for the compute)."""

    def __init__(self, backend, world_n, group=None, library_comm=False):
        self.backend = backend
        self.world_n = int(world_n)
        self.group = group
        self.library_comm = bool(library_comm)

    def eval(self, theta, U, score):
        if self.library_comm:
            return self.backend.fitc_eval_sharded(theta, U, score, self.world_n)
        return self.backend.fitc_eval_sharded(theta, U, score, self.world_n,
                                              lambda t: allreduce_sum_(t, self.group))


def sharded_metrics(local_sums, n_total, group=None):
    """Finish KF:276-292 from per-rank raw sums (see gps_test_metrics): returns the metric dict."""
    t = torch.as_tensor(np.asarray(local_sums, dtype=np.float64)).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl":
        t = t.cuda()
    allreduce_sum_(t, group)
    s = t.cpu().numpy()
    n = float(n_total)
    return {"mse": s[0] / n, "smse": s[0] / s[1], "logs": s[2] / n, "crps": s[3] / n,
            "msll": (s[2] - s[4]) / n, "coverage": s[5] / n}


def sharded_predict_metrics(ctx, theta, Xs, ys, inducing_x=None, group=None):
    """Prediction + test scoring split by rows of the test set: every rank predicts its
    contiguous block of test rows with its own context (which holds the full training set; for FITC the
    M x M factors are replicated) and the six metric sums are all-reduced.  Returns the metric dict."""
    rank, size = world()
    t = int(Xs.shape[0])
    lo, hi = row_block(t, rank, size)
    if hi > lo:
        if inducing_x is None:
            mean, var = ctx.full_predict(theta, Xs[lo:hi])
        else:
            mean, var = ctx.fitc_predict(theta, inducing_x, Xs[lo:hi])
        _, sums = ctx.test_metrics(mean, var, ys[lo:hi], return_sums=True)
    else:
        sums = np.zeros(6)
    return sharded_metrics(sums, t, group)


def sharded_grid(eval_fn, ls, sd, group=None, device=None):
    """Evaluate eval_fn(ls_subset, sd_subset) -> values on this rank's round-robin share of the grid
    and combine: every rank returns the full vector (the input of CP:114's `matrix(..., nrow=50)`)."""
    rank, size = world()
    ls, sd = np.asarray(ls, dtype=np.float64).ravel(), np.asarray(sd, dtype=np.float64).ravel()
    idx = round_robin(ls.size, rank, size)
    full = np.zeros(ls.size)
    if idx.size:
        full[idx] = np.asarray(eval_fn(ls[idx], sd[idx]), dtype=np.float64)
    t = torch.from_numpy(full)
    if device is not None:
        t = t.to(device)
    allreduce_sum_(t, group)
    return t.cpu().numpy()


def grid_matrix(values, n_l, n_noise):
    """Reshape grid results the way CP:113-114 does: `sapply(l_range, function(xx) mapply(f, xx,
    noise_range))` then `matrix(res, nrow = 50)` — rows index the noise s.d., columns the length
    scale.  `values` is ordered length-scale-major (all noise values for l_0, then l_1, ...)."""
    return np.asarray(values).reshape(n_l, n_noise).T
