"""A/B of capped persistent grids for the launches that share the GPU with POTRF's diagonal-block chain (knob 11: the
overlapped TRTRI merges, knob 12: the trailing updates; value = number of CTAs, 0 = one CTA per slice as before; the
GPU holds 296 of these CTAs): ms per headline evaluation, and the objective (must not change in any bit)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpscore_b200 import api, synth
ctx = api.Context(0)
s = torch.cuda.Stream(); ctx.set_stream(s)
X, y = synth.kin40k_like(10000); theta = synth.hyper_point("P1")
ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
combos = [(0, 0), (280, 0), (264, 0), (232, 0), (0, 280), (0, 264), (280, 280), (264, 264), (232, 264)]
if len(sys.argv) > 1:
    combos = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for ct, cl in combos:
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 11, ct))
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 12, cl))
    for _ in range(2): v, g = ctx.full_eval(theta, "crps")
    e0.record(s)
    for _ in range(5): v, g = ctx.full_eval(theta, "crps")
    e1.record(s); s.synchronize()
    st = ctx.last_stage_ms()
    print("cap trtri %3d trail %3d: %.2f ms/eval  potrf+trtri %.2f  obj %.17g |g| %.17g" % (
        ct, cl, e0.elapsed_time(e1) / 5, st["potrf"] + st["trtri"], v, float(np.linalg.norm(g))), flush=True)
