"""Smallest program that exercises the headline path: a few full-GP LOO-CRPS obj+grad evaluations at
N rows (default 10000) and FITC M=20 evaluations.  Used under ncu (launch list / --set full)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpscore_b200 import api, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
X, y = synth.kin40k_like(N)
theta = synth.hyper_point("P1")
ctx = api.Context(0)
ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
for i in range(reps):
    t0 = time.perf_counter()
    v, g = ctx.full_eval(theta, "crps")
    print("full eval %d: %.3f ms obj %.12g" % (i, (time.perf_counter() - t0) * 1e3, v))
U = synth.inducing_init(20)
for i in range(reps):
    t0 = time.perf_counter()
    v, g, gU = ctx.fitc_eval(theta, U, "crps")
    print("fitc eval %d: %.3f ms obj %.12g" % (i, (time.perf_counter() - t0) * 1e3, v))
print("launches", ctx.launch_count())
