"""A few FITC M=20 evaluations at N rows (default 1e6) for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpscore_b200 import api, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = api.Context(0)
ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 2, variant))
X, y = synth.kin40k_like(N, seed=7)
ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
U = synth.inducing_init(20)
theta = synth.hyper_point("P1")
for _ in range(3):
    print(ctx.fitc_eval(theta, U, "crps")[0])
