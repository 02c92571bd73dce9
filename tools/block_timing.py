"""FITC block objectives (4-fold DSS, kc) against the LOO CRPS at the same shape: ms per objective + gradient evaluation
(CUDA events around gps_fitc_eval, host arguments).  M = 20 compares the row kernels of gps_fitc.cu with the matrix
form (switch-over lowered); larger M is the matrix form."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpscore_b200 import api, synth

ctx = api.Context(0)
stream = torch.cuda.Stream(); ctx.set_stream(stream)
theta = synth.hyper_point("P1")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def timed(fn, reps):
    fn(); fn()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps): out = fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

for n, ms in ((10000, (20, 64, 256)), (1000000, (20, 64, 256, 1024))):
    X, y = synth.kin40k_like(n, seed=7)
    ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
    rng = np.random.default_rng(1)
    for m in ms:
        U = X[rng.choice(n, m, replace=False)] + 0.05 * rng.standard_normal((m, 8))
        reps = 3 if m >= 1024 else (10 if n > 100000 else 30)
        line = []
        for forced in ((False, True) if m <= 32 else (False,)):
            ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 3, 1 if forced else 33))
            for score in ("crps", "dss", "kc"):
                t, out = timed(lambda: ctx.fitc_eval(theta, U, score), reps)
                line.append("%s%s %.3f ms" % (score, " (matrix form)" if forced else "", t))
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 3, 33))
        print("N=%d M=%d: %s" % (n, m, " | ".join(line)), flush=True)
ctx.close()
