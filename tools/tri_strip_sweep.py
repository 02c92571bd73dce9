"""A/B of the strip policy of the TRTRI merges issued behind POTRF (knob 10: 0 = normal 64-row CTAs, 32 / 16 = four /
eight shorter CTAs per tile, which hold an SM for less time when a chain kernel is waiting for a slot), crossed with the
node split (knob 9): ms per headline evaluation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpscore_b200 import api, synth
ctx = api.Context(0)
s = torch.cuda.Stream(); ctx.set_stream(s)
X, y = synth.kin40k_like(10000); theta = synth.hyper_point("P1")
ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for pct in (50, 66):
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 9, pct))
    for strip in (0, 32, 16):
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 10, strip))
        for _ in range(2): v, g = ctx.full_eval(theta, "crps")
        e0.record(s)
        for _ in range(5): v, g = ctx.full_eval(theta, "crps")
        e1.record(s); s.synchronize()
        st = ctx.last_stage_ms()
        print("split %d%% tri_strip %2d: %.2f ms/eval  potrf+trtri %.2f  obj %.15g" % (
            pct, strip, e0.elapsed_time(e1) / 5, st["potrf"] + st["trtri"], v), flush=True)
