"""A/B of the POTRF / TRTRI overlap (gps_dbg_set_variant knob 4) on the headline evaluation."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpscore_b200 import api, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
X, y = synth.kin40k_like(N)
theta = synth.hyper_point("P1")
ctx = api.Context(0)
ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
for mode in (2, 0, 1, 1):
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 4, mode))
    for _ in range(2):
        v, g = ctx.full_eval(theta, "crps")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.full_eval(theta, "crps")
    dt = (time.perf_counter() - t0) / 5
    st = ctx.last_stage_ms()
    print("overlap=%d  %.2f ms/eval  obj %.15g  |g| %.12g  stages %s" % (
        mode, dt * 1e3, v, float((g * g).sum() ** 0.5), {k: round(x, 2) for k, x in st.items()}))
