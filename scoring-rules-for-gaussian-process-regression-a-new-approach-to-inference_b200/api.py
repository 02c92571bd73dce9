"""Host-side mirror of the reference's function names over the CUDA library.

The reference has no importable module: its hot path is loop-body code in four scripts
(SURVEY.md §0).  This module keeps the *names and positional signatures* those loop bodies use —
`ARD`, `chol_solve`, `Q`, `crps`, `logs`, `cal_mean_and_cov`, `spgp_cal_mean_and_cov`, `SMSE`,
`trivial_loss` — and adds the fused entry points the loops are rewritten onto:

    CRPS_ave = full_loo_objective(train_x, train_y, para_k, para_l, para_noise, "crps")   # KF:239-245
    CRPS_ave.backward()                                                                    # KF:252
    with torch.no_grad(): para_l -= lr * para_l.grad ...                                   # KF:254-260 unchanged

torch tensors are buffers only: every number is produced by libgpscore.so through ctypes
(lib.py) — including the matrix products inside the twins `Q`, `cal_mean_and_cov` and
`spgp_cal_mean_and_cov` (`gps_matmul`).

Stream ordering: unless a stream is pinned with `Context.set_stream`, every call first points the
library at torch's CURRENT stream on the context's device, so tensors produced by torch ops
(`.double()`, `.cuda()`, NCCL all-reduces) and the library's kernels are ordered like any two torch
ops would be.  Hyper-parameterisation as in the reference: para_k = log sf^2, para_l = log l
(one element or [1, D]), para_noise = log sn^2 (KF:7-12, KF:239).
"""
import ctypes as C
import math
import weakref

import numpy as np
import torch

from . import lib as _L

JITTER = 1e-3  # K20:36

# Module globals the reference's helpers read (KF:33-36, KF:44, KF:122).  Set them like the
# scripts do (`gp.sigma_noise_sq = torch.exp(para_noise)`) or pass the keyword overrides.
para_k = None
para_l = None
sigma_noise_sq = None


def _as_f64(a, device=None):
    if isinstance(a, torch.Tensor):
        t = a.detach()
        if t.dtype != torch.float64:
            t = t.double()
        if device is not None and t.device != device:
            t = t.to(device)
        return t.contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=device)


def _host_vec(a):
    if isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous:
        return a.reshape(-1)
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().double().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())


def _dp(arr):
    return arr.ctypes.data_as(C.POINTER(C.c_double))


def _scalar(v):
    if isinstance(v, torch.Tensor):
        return float(v.detach().reshape(-1)[0])
    return float(np.asarray(v).reshape(-1)[0])


class _IO:
    """Host-side staging of one evaluation's small arguments (theta, U -> objective, gradients) with the ctypes
    pointers made once: building five `ndarray.ctypes.data_as(...)` pointers and three fresh arrays per call costs
    more host time than the three kernel launches of a FITC evaluation."""

    def __init__(self, D, M):
        self.th = np.zeros(D + 2)
        self.U = np.zeros(max(M * D, 1))
        self.obj = np.zeros(1)
        self.g = np.zeros(D + 2)
        self.gU = np.zeros(max(M * D, 1))
        self.p_th, self.p_U, self.p_obj, self.p_g, self.p_gU = (_dp(a) for a in (self.th, self.U, self.obj, self.g, self.gU))


class Context:
    """One gps_ctx on one GPU.  Owns the padded copy of the training set and all workspaces."""

    def __init__(self, device=None):
        self._lib = _L.load()
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        if isinstance(device, torch.device):
            device = device.index or 0
        h = C.c_void_p()
        code = self._lib.gps_create(int(device), C.byref(h))
        if code != _L.GPS_OK:
            _L.check(None, code)
        self._h = h
        self.device = torch.device("cuda", int(device))
        self.N = 0
        self.D = 0
        self._data_key = None
        self._pinned = False        # set_stream(stream) pins; otherwise torch's current stream is followed
        self._cur_stream = None
        self._io_cache = {}

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gps_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, code):
        _L.check(self._h, code)

    def _enter(self):
        """Order the library behind torch: enqueue on torch's current stream of this device (the
        legacy default stream is passed as cudaStreamLegacy = 1; NULL would select the context's own)."""
        if self._pinned:
            return
        s = torch.cuda.current_stream(self.device).cuda_stream or 1
        if s != self._cur_stream:
            self._check(self._lib.gps_set_stream(self._h, s))
            self._cur_stream = s

    # ---- data --------------------------------------------------------------------------------
    def set_data(self, X, y):
        """train_x [N, D], train_y [N] or [N, 1] (KF:208-209); host or device."""
        self._enter()
        X = _as_f64(X)
        y = _as_f64(y).reshape(-1)
        if X.dim() != 2 or y.numel() != X.shape[0]:
            raise ValueError("set_data: X must be [N, D] and y must have N elements")
        self._check(self._lib.gps_set_data(self._h, X.data_ptr(), y.data_ptr(), X.shape[0], X.shape[1]))
        self.N, self.D = int(X.shape[0]), int(X.shape[1])
        self._data_key = None       # whatever _bind_data cached no longer describes the context
        self._y_train = y           # mean / unbiased variance of the targets (KF:113-114) are formed on first use
        self._y_stats = None

    def _io(self, M):
        key = (self.D, M)
        io = self._io_cache.get(key)
        if io is None:
            io = self._io_cache[key] = _IO(self.D, M)
        return io

    def _theta(self, theta):
        th = _host_vec(theta)
        if th.size == 3 and self.D > 1:  # isotropic 1-element para_l (K20:422): broadcast (KF:8)
            th = np.concatenate([[th[0]], np.full(self.D, th[1]), [th[2]]])
        if th.size != self.D + 2:
            raise ValueError("theta must have D + 2 = %d elements" % (self.D + 2))
        return th

    # ---- full GP -----------------------------------------------------------------------------
    def full_eval(self, theta, score, grad=True):
        """Objective and gradient wrt theta = [a, b_1..b_D, c] (KF:239-252 / 329-339 / 416-428)."""
        self._enter()
        io = self._io(0)
        io.th[:] = self._theta(theta)
        sc = _L.SCORES[score] if isinstance(score, str) else int(score)
        self._check(self._lib.gps_full_eval(self._h, io.p_th, sc, io.p_obj, io.p_g if grad else None))
        return float(io.obj[0]), (io.g.copy() if grad else None)

    def full_descend(self, theta, score, lr, iters):
        """`iters` steps of the scripts' fixed-step gradient descent (KF:237-260) in one call.
        Returns (theta after the last step, objective before each step)."""
        self._enter()
        th = self._theta(theta).copy()
        trace = np.zeros(int(iters))
        sc = _L.SCORES[score] if isinstance(score, str) else int(score)
        self._check(self._lib.gps_full_descend(self._h, _dp(th), sc, float(lr), int(iters), _dp(trace)))
        return th, trace

    def fitc_descend(self, theta, U, score, lr, lr_u, iters, jitter=JITTER):
        """K20:219-251 in one call: theta -= lr * grad, inducing_x -= lr_u * grad (K20:326-327)."""
        self._enter()
        th = self._theta(theta).copy()
        Uh = _host_vec(U).copy()
        M = Uh.size // self.D
        trace = np.zeros(int(iters))
        sc = _L.SCORES[score] if isinstance(score, str) else int(score)
        self._check(self._lib.gps_fitc_descend(self._h, _dp(th), _dp(Uh), M, float(jitter), sc, float(lr), float(lr_u),
                                               int(iters), _dp(trace)))
        return th, Uh.reshape(M, self.D), trace

    def full_loo(self):
        """mean_term, cov_term of KF:243-244 for the last crps/logs evaluation, as [N, 1] tensors."""
        self._enter()
        m = torch.empty(self.N, dtype=torch.float64, device=self.device)
        v = torch.empty(self.N, dtype=torch.float64, device=self.device)
        self._check(self._lib.gps_full_loo(self._h, m.data_ptr(), v.data_ptr()))
        return m.view(-1, 1), v.view(-1, 1)

    def full_predict(self, theta, Xs):
        """Predictive mean and variance (diagonal of KF:121-126's covariance) at Xs [T, D]."""
        self._enter()
        th = self._theta(theta)
        Xs = _as_f64(Xs)
        T = int(Xs.shape[0])
        m = torch.empty(T, dtype=torch.float64, device=self.device)
        v = torch.empty(T, dtype=torch.float64, device=self.device)
        self._check(self._lib.gps_full_predict(self._h, _dp(th), Xs.data_ptr(), T, m.data_ptr(), v.data_ptr()))
        return m.view(-1, 1), v.view(-1, 1)

    # ---- FITC --------------------------------------------------------------------------------
    def fitc_eval(self, theta, U, score, jitter=JITTER):
        """Objective and gradients (theta, inducing inputs) of K20:222-236 / 329-344 / 434-452."""
        self._enter()
        Uh = _host_vec(U)
        M = Uh.size // self.D
        io = self._io(M)
        io.th[:] = self._theta(theta)
        io.U[:] = Uh
        sc = _L.SCORES[score] if isinstance(score, str) else int(score)
        self._check(self._lib.gps_fitc_eval(self._h, io.p_th, io.p_U, M, float(jitter), sc, io.p_obj, io.p_g, io.p_gU))
        return float(io.obj[0]), io.g.copy(), io.gU.copy().reshape(M, self.D)

    def fitc_acc_len(self, M):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        code = self._lib.gps_fitc_acc_len(int(M), self.D, C.byref(a), C.byref(b), C.byref(c))
        if code != _L.GPS_OK:
            raise _L.GpsError(code, "fitc_acc_len")
        return a.value, b.value, c.value

    # ---- row-sharded FITC (one process per GPU) ----------------------------------------------
    def comm_init(self, group=None):
        """Create this context's NCCL communicator inside the library.  Rank 0 draws the unique id and
        torch.distributed (any backend) carries its 128 bytes to the other ranks; after that the
        all-reduces of `fitc_eval_sharded` are issued by libgpscore itself on the context's stream."""
        import torch.distributed as dist
        self._enter()
        try:    # bind the libnccl this process already uses (torch's bundled copy), not a second one
            import nvidia.nccl
            import os
            for base in list(getattr(nvidia.nccl, "__path__", [])):
                cand = os.path.join(base, "lib", "libnccl.so.2")
                if os.path.isfile(cand):
                    self._lib.gps_comm_set_library(cand.encode())
                    break
        except Exception:
            pass
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        uid = (C.c_char * 128)()
        if rank == 0:
            code = self._lib.gps_comm_unique_id(uid)
            if code != _L.GPS_OK:
                raise _L.GpsError(code, "gps_comm_unique_id failed (libnccl.so.2 not loadable?)")
        backend = dist.get_backend(group)
        t = torch.frombuffer(bytearray(uid.raw), dtype=torch.uint8).clone()
        if backend == "nccl":
            t = t.to(self.device)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(t.cpu().numpy().tobytes())
        self._check(self._lib.gps_comm_init(self._h, C.c_char_p(raw), rank, world))
        self.rank, self.world = rank, world

    def comm_transport(self, peer_memory=True):
        """Select how the M <= 31 row-sharded evaluation exchanges its accumulators: inside the pass kernels over
        peer memory (default) or with ncclAllReduce between them.  Returns True if peer memory is in effect."""
        act = C.c_int()
        self._check(self._lib.gps_comm_set_transport(self._h, 1 if peer_memory else 0, C.byref(act)))
        return bool(act.value)

    def comm_allreduce_(self, t):
        """In-place sum of a float64 device tensor over the context's communicator (library NCCL)."""
        self._enter()
        assert t.dtype == torch.float64 and t.is_cuda and t.is_contiguous()
        self._check(self._lib.gps_comm_allreduce_sum(self._h, t.data_ptr(), t.numel()))
        return t

    def fitc_eval_sharded(self, theta, U, score, world_n, allreduce=None, jitter=JITTER):
        """Row-sharded FITC evaluation: this context holds a contiguous block of the `world_n` rows.
        allreduce=None (production): gps_fitc_eval_sharded — the three all-reduces run inside the
        library over its NCCL communicator (`comm_init`), nothing synchronises the host between the
        passes.  A callable (`allreduce(tensor)` sums a 1-D float64 device tensor in place across
        ranks) selects the staged protocol of include/gpscore.h instead, for callers that bring their
        own collective.  The block objectives "dss" / "kc" (world_n % 4 == 0) shard through
        the library communicator only (allreduce=None); the ranks must hold consecutive row blocks in rank
        order (`dist.row_block`), because the four folds are quarters of the global row order."""
        self._enter()
        th = self._theta(theta)
        Uh = _host_vec(U)
        M = Uh.size // self.D
        sc = _L.SCORES[score] if isinstance(score, str) else int(score)
        obj = np.zeros(1)
        g = np.zeros(self.D + 2)
        gU = np.zeros(M * self.D)
        if allreduce is None:
            self._check(self._lib.gps_fitc_eval_sharded(self._h, _dp(th), _dp(Uh), M, float(jitter), sc, int(world_n),
                                                        _dp(obj), _dp(g), _dp(gU)))
            return float(obj[0]), g, gU.reshape(M, self.D)
        l1, l2, l3 = self.fitc_acc_len(M)
        key = (M, self.D)
        if getattr(self, "_acc_key", None) != key:
            self._acc = [torch.zeros(n, dtype=torch.float64, device=self.device) for n in (l1, l2, l3)]
            self._acc_key = key
        a1, a2, a3 = self._acc
        self._check(self._lib.gps_fitc_begin(self._h, _dp(th), _dp(Uh), M, float(jitter), sc, int(world_n)))
        self._check(self._lib.gps_fitc_pass1(self._h, a1.data_ptr()))
        allreduce(a1)
        self._after_collective()
        self._check(self._lib.gps_fitc_pass2(self._h, a1.data_ptr(), a2.data_ptr()))
        allreduce(a2)
        self._after_collective()
        self._check(self._lib.gps_fitc_pass3(self._h, a2.data_ptr(), a3.data_ptr()))
        allreduce(a3)
        self._after_collective()
        self._check(self._lib.gps_fitc_finish(self._h, a2.data_ptr(), a3.data_ptr(), _dp(obj), _dp(g), _dp(gU)))
        return float(obj[0]), g, gU.reshape(M, self.D)

    def fitc_descend_sharded(self, theta, U, score, world_n, lr, lr_u, iters, jitter=JITTER):
        """`fitc_descend` on a row-sharded problem (after `comm_init`): every rank returns the same
        (theta, U, objective trace)."""
        self._enter()
        th = self._theta(theta).copy()
        Uh = _host_vec(U).copy()
        M = Uh.size // self.D
        trace = np.zeros(int(iters))
        sc = _L.SCORES[score] if isinstance(score, str) else int(score)
        self._check(self._lib.gps_fitc_descend_sharded(self._h, _dp(th), _dp(Uh), M, float(jitter), sc, int(world_n),
                                                       float(lr), float(lr_u), int(iters), _dp(trace)))
        return th, Uh.reshape(M, self.D), trace

    def _after_collective(self):
        """The caller's collective ran on torch's current stream.  Following that stream keeps the
        next pass ordered behind it; a context pinned to another stream has to wait for it."""
        if self._pinned:
            torch.cuda.current_stream(self.device).synchronize()
        else:
            self._enter()

    def fitc_loo(self):
        self._enter()
        m = torch.empty(self.N, dtype=torch.float64, device=self.device)
        v = torch.empty(self.N, dtype=torch.float64, device=self.device)
        self._check(self._lib.gps_fitc_loo(self._h, m.data_ptr(), v.data_ptr()))
        return m.view(-1, 1), v.view(-1, 1)

    def fitc_predict(self, theta, U, Xs, jitter=JITTER, score="nlml", force_matrix_form=False):
        """Predictive mean / variance of K20:270-277 (diagonal only) at Xs."""
        self._enter()
        th = self._theta(theta)
        Uh = _host_vec(U)
        M = Uh.size // self.D
        # an objective-only evaluation leaves the factors (L_A, L_C, beta) in place on every path
        obj = np.zeros(1)
        self._check(self._lib.gps_fitc_eval(self._h, _dp(th), _dp(Uh), M, float(jitter), _L.SCORES[score],
                                            _dp(obj), None, None))
        Xs = _as_f64(Xs)
        T = int(Xs.shape[0])
        m = torch.empty(T, dtype=torch.float64, device=self.device)
        v = torch.empty(T, dtype=torch.float64, device=self.device)
        self._check(self._lib.gps_fitc_predict(self._h, Xs.data_ptr(), T, m.data_ptr(), v.data_ptr()))
        return m.view(-1, 1), v.view(-1, 1)

    # ---- metrics / element-wise twins ---------------------------------------------------------
    def test_metrics(self, mean, var, y, y_train=None, return_sums=False):
        """mse, SMSE (KF:128-134), logs, crps, MSLL (KF:110-119), +-2 sd coverage (KF:288-292)."""
        self._enter()
        mean, var, y = _as_f64(mean).reshape(-1), _as_f64(var).reshape(-1), _as_f64(y).reshape(-1)
        if y_train is not None:
            yt = _as_f64(y_train).reshape(-1)
            ytm, ytv = float(yt.mean()), float(yt.var(unbiased=True))
        else:
            if self._y_stats is None:
                yt = self._y_train
                self._y_stats = (float(yt.mean()), float(yt.var(unbiased=True)) if yt.numel() > 1 else 0.0)
            ytm, ytv = self._y_stats
        out = np.zeros(12)
        self._check(self._lib.gps_test_metrics(self._h, mean.data_ptr(), var.data_ptr(), y.data_ptr(), mean.numel(),
                                               ytm, ytv, _dp(out)))
        res = dict(zip(("mse", "smse", "logs", "crps", "msll", "coverage"), out[:6].tolist()))
        if return_sums:
            return res, out[6:].copy()
        return res

    def ard(self, x, xp, a, b):
        self._enter()
        x, xp = _as_f64(x), _as_f64(xp)
        bh = _host_vec(b)
        out = torch.empty(x.shape[0], xp.shape[0], dtype=torch.float64, device=x.device)
        self._check(self._lib.gps_ard(self._h, x.data_ptr(), x.shape[0], xp.data_ptr(), xp.shape[0], x.shape[1],
                                      _scalar(a), _dp(bh), bh.size, out.data_ptr()))
        return out

    def chol_solve(self, B, A):
        self._enter()
        B2, A2 = _as_f64(B), _as_f64(A)
        n = A2.shape[0]
        B2 = B2.reshape(n, -1)
        out = torch.empty_like(B2)
        self._check(self._lib.gps_chol_solve(self._h, B2.data_ptr(), A2.data_ptr(), n, B2.shape[1], out.data_ptr()))
        return out

    def score(self, m, c, y, which):
        self._enter()
        m, c, y = _as_f64(m).reshape(-1), _as_f64(c).reshape(-1), _as_f64(y).reshape(-1)
        out = np.zeros(1)
        self._check(self._lib.gps_score(self._h, m.data_ptr(), c.data_ptr(), y.data_ptr(), m.numel(),
                                        _L.SCORES[which], _dp(out)))
        return float(out[0])

    def grid_eval(self, x, y, ls, noise_sd, which):
        """CP:109-144: objective `which` at every (length scale, noise s.d.) pair."""
        self._enter()
        x, y = _as_f64(x).reshape(-1), _as_f64(y).reshape(-1)
        ls, sd = _host_vec(ls), _host_vec(noise_sd)
        out = np.zeros(ls.size)
        self._check(self._lib.gps_grid_eval(self._h, x.data_ptr(), y.data_ptr(), x.numel(), _dp(ls), _dp(sd), ls.size,
                                            _L.GRID_KINDS[which], _dp(out)))
        return out

    # ---- accounting --------------------------------------------------------------------------
    def set_stream(self, stream=None):
        """Pin the context to a torch stream (so torch.cuda.Event on it brackets the work); None
        un-pins: the context follows torch's current stream again."""
        if stream is None:
            self._pinned = False
            self._cur_stream = None
            return
        self._check(self._lib.gps_set_stream(self._h, stream.cuda_stream or 1))
        self._pinned = True
        self._cur_stream = stream.cuda_stream or 1

    def matmul(self, A, B):
        """torch.mm twin on the library's tile GEMM (gps_matmul)."""
        self._enter()
        A2, B2 = _as_f64(A), _as_f64(B)
        if A2.dim() != 2 or B2.dim() != 2 or A2.shape[1] != B2.shape[0]:
            raise ValueError("matmul: shapes %s x %s" % (tuple(A2.shape), tuple(B2.shape)))
        out = torch.empty(A2.shape[0], B2.shape[1], dtype=torch.float64, device=A2.device)
        self._check(self._lib.gps_matmul(self._h, A2.data_ptr(), B2.data_ptr(), A2.shape[0], A2.shape[1], B2.shape[1],
                                         out.data_ptr()))
        return out

    def set_gemm_timing(self, on):
        self._check(self._lib.gps_set_gemm_timing(self._h, 1 if on else 0))

    def launch_floor_us(self, launches=3, reps=50):
        """Wall-clock floor of an evaluation call: `launches` empty kernels + one stream synchronisation."""
        self._enter()
        us = C.c_double()
        self._check(self._lib.gps_dbg_launch_floor(self._h, int(launches), int(reps), C.byref(us)))
        return us.value

    def launch_count(self):
        return int(self._lib.gps_launch_count(self._h))

    def last_stage_ms(self):
        """Device time per stage of the last CRPS/LOGS obj+grad evaluation (CUDA events)."""
        ms = np.zeros(7)
        self._check(self._lib.gps_last_stage_ms(self._h, _dp(ms)))
        return dict(zip(("gram", "potrf", "trtri", "lauum", "score", "symprod", "contract"), ms.tolist()))

    def last_gemm_ms(self):
        ms, n = C.c_double(), C.c_int64()
        self._lib.gps_last_gemm_ms(self._h, C.byref(ms), C.byref(n))
        return ms.value, n.value


_default = {}


def default_context(device=None):
    """Process-wide context per GPU, created on first use."""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if isinstance(device, torch.device):
        device = device.index or 0
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]


def _bind_data(ctx, X, y):
    """Upload the training set unless THESE tensor objects, unmodified, are what the context holds.
    Identity is by weak reference (a new tensor that the caching allocator places at a freed tensor's
    address is a different object), modification by torch's version counters; anything else —
    numpy arrays, tensors whose objects died — is uploaded again: N D doubles are nothing next to
    the evaluation."""
    if isinstance(X, torch.Tensor) and isinstance(y, torch.Tensor):
        k = ctx._data_key
        if (k is not None and k[0]() is X and k[1]() is y and k[2] == (X._version, y._version, tuple(X.shape), tuple(y.shape))):
            return
        ctx.set_data(X, y)
        ctx._data_key = (weakref.ref(X), weakref.ref(y), (X._version, y._version, tuple(X.shape), tuple(y.shape)))
    else:
        ctx.set_data(X, y)


def _theta_from_leaves(pk, pl, pn, D):
    b = _host_vec(pl)
    if b.size == 1:
        b = np.full(D, b[0])
    return np.concatenate([[_scalar(pk)], b, [_scalar(pn)]])


class _FusedObjective(torch.autograd.Function):
    """forward = one C-ABI call returning value AND gradient; backward scales the cached gradient.
    Makes `obj.backward()` populate `.grad` of para_k / para_l / para_noise / inducing_x exactly as
    the scripts' loops expect (KF:252-260, K20:236-251)."""

    @staticmethod
    def forward(fctx, pk, pl, pn, U, X, y, score, jitter, ctx):
        D = X.shape[1]
        theta = _theta_from_leaves(pk, pl, pn, D)
        _bind_data(ctx, X, y)
        if U is None:
            val, g = ctx.full_eval(theta, score)
            gU = None
        else:
            val, g, gU = ctx.fitc_eval(theta, U, score, jitter)
        fctx.grads = (g, gU)
        fctx.meta = (pk, pl, pn, U)
        out = torch.tensor(val, dtype=pk.dtype, device=pk.device)
        if score in ("nlml", "dss"):
            out = out.reshape(1, 1)  # Neg_logL / dss_ave come out of a [1, 1] product in the scripts (KF:334, KF:107)
        return out

    @staticmethod
    def backward(fctx, gout):
        g, gU = fctx.grads
        pk, pl, pn, U = fctx.meta
        s = float(gout.reshape(-1)[0])
        gk = torch.tensor([g[0] * s], dtype=pk.dtype, device=pk.device).reshape(pk.shape)
        gb = g[1:-1]
        if pl.numel() == 1:
            gl = torch.tensor([gb.sum() * s], dtype=pl.dtype, device=pl.device).reshape(pl.shape)
        else:
            gl = torch.tensor(gb * s, dtype=pl.dtype, device=pl.device).reshape(pl.shape)
        gn = torch.tensor([g[-1] * s], dtype=pn.dtype, device=pn.device).reshape(pn.shape)
        gu = None
        if U is not None:
            gu = torch.tensor(gU * s, dtype=U.dtype, device=U.device).reshape(U.shape)
        return gk, gl, gn, gu, None, None, None, None, None


def full_loo_objective(train_x, train_y, para_k, para_l, para_noise, score="crps", ctx=None):
    """The six statements KF:239-245 (crps), KF:416-424 (logs) or KF:329-334 (nlml) as one fused,
    differentiable call.  Returns the scalar the scripts call CRPS_ave / logs_ave / Neg_logL."""
    ctx = ctx or default_context()
    return _FusedObjective.apply(para_k, para_l, para_noise, None, train_x, train_y, score, JITTER, ctx)


def fitc_loo_objective(train_x, train_y, inducing_x, para_k, para_l, para_noise, score="crps", jitter=JITTER,
                       ctx=None):
    """K20:222-234 (crps), K20:434-447 (logs) or K20:329-340 (nlml) as one fused, differentiable call;
    gradients flow to inducing_x as well (K20:247)."""
    ctx = ctx or default_context()
    return _FusedObjective.apply(para_k, para_l, para_noise, inducing_x, train_x, train_y, score, jitter, ctx)


def predict_diag(train_x, train_y, test_x, para_k, para_l, para_noise, inducing_x=None, jitter=JITTER, ctx=None):
    """KF:267-273 / K20:270-277 with only the diagonal of the covariance formed: (mean, var) [T, 1]."""
    ctx = ctx or default_context()
    _bind_data(ctx, train_x, train_y)
    theta = _theta_from_leaves(para_k, para_l, para_noise, train_x.shape[1])
    if inducing_x is None:
        return ctx.full_predict(theta, test_x)
    return ctx.fitc_predict(theta, inducing_x, test_x, jitter)


# ---- same-signature twins of the reference's helpers ----------------------------------------------
def ARD(x, xp, a, b):
    """KF:7-23."""
    return default_context().ard(x, xp, a, b)


def chol_solve(B, A):
    """KF:25-29: A^-1 B."""
    out = default_context().chol_solve(B, A)
    return out.reshape(B.shape) if isinstance(B, torch.Tensor) else out


def Q(a, u, b, para_k=None, para_l=None, jitter=JITTER):
    """KF:32-39: K_au (K_uu + 1e-3 I)^-1 K_ub; para_k / para_l default to the module globals."""
    g = globals()
    pk = g["para_k"] if para_k is None else para_k
    pl = g["para_l"] if para_l is None else para_l
    K_au = ARD(a, u, pk, pl)
    K_uu = ARD(u, u, pk, pl)
    K_uu = K_uu + jitter * torch.eye(K_uu.shape[0], dtype=K_uu.dtype, device=K_uu.device)
    K_ub = ARD(u, b, pk, pl)
    return default_context().matmul(K_au, chol_solve(K_ub, K_uu))


def crps(m, c, data_y):
    """KF:60-68; c is the variance."""
    return torch.tensor(default_context().score(m, c, data_y, "crps"), dtype=torch.float64)


def logs(m, c, data_y):
    """KF:52-57."""
    return torch.tensor(default_context().score(m, c, data_y, "logs"), dtype=torch.float64)


def _noise(sn2):
    v = globals()["sigma_noise_sq"] if sn2 is None else sn2
    if v is None:
        raise ValueError("sigma_noise_sq is not set (module global, KF:122) and no override was given")
    return _scalar(v)


def cal_mean_and_cov(k1, k2, k3, num, eye_num, data_y, sigma_noise_sq=None):
    """KF:121-126 (full T x T covariance, as the reference returns it)."""
    sn2 = _noise(sigma_noise_sq)
    eye = torch.eye(eye_num, dtype=k2.dtype, device=k2.device)
    Kn = k2 + sn2 * eye
    sol = chol_solve(torch.cat([data_y.reshape(eye_num, -1), k1.t()], dim=1), Kn)
    prod = default_context().matmul(k1, sol)      # [k1 Kn^-1 y | k1 Kn^-1 k1'] in one product
    res_mean = prod[:, :1]
    res_cov = sn2 * torch.eye(num, dtype=k2.dtype, device=k2.device) + k3 - prod[:, 1:]
    return res_mean, res_cov


def spgp_cal_mean_and_cov(k1, Q1, Q2, k2, num_test, num_jitter, data_y, sigma_noise_sq=None):
    """K20:76-83."""
    sn2 = _noise(sigma_noise_sq)
    G = torch.diag(torch.diag(k1 - Q1) + sn2)
    sol = chol_solve(torch.cat([data_y.reshape(num_jitter, -1), Q2.t()], dim=1), Q1 + G)
    prod = default_context().matmul(Q2, sol)
    mean_term = prod[:, :1]
    cov_term = sn2 * torch.eye(num_test, dtype=k2.dtype, device=k2.device) + k2 - prod[:, 1:]
    return mean_term, cov_term


def SMSE(m, data_y, data_yp):
    """KF:128-134."""
    r = default_context().test_metrics(m, torch.ones_like(_as_f64(m)), data_y, data_yp)
    return torch.tensor(r["smse"], dtype=torch.float64)


def trivial_loss(m, c, data_y, data_yp):
    """KF:110-119 (MSLL)."""
    r = default_context().test_metrics(m, c, data_y, data_yp)
    return torch.tensor(r["msll"], dtype=torch.float64)


def test_metrics(m, c, data_y, data_yp):
    """KF:276-292 in one call."""
    return default_context().test_metrics(m, c, data_y, data_yp)
