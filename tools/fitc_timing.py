"""FITC M=20 evaluation time at N = 1e4 and 1e6 for both row-pass formulations."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpscore_b200 import api, synth  # noqa: E402

ctx = api.Context(0)
U = synth.inducing_init(20)
theta = synth.hyper_point("P1")
for N in (10000, 1000000):
    X, y = synth.kin40k_like(N, seed=7)
    ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
    for variant in (0, 1):
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 2, variant))
        for _ in range(3):
            v = ctx.fitc_eval(theta, U, "crps")[0]
        torch.cuda.synchronize()
        reps = 200 if N == 10000 else 20
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.fitc_eval(theta, U, "crps")
        dt = (time.perf_counter() - t0) / reps
        print("N=%8d variant %d: %.3f ms/eval  obj %.12g" % (N, variant, dt * 1e3, v))
