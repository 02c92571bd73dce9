"""A/B of knob 13 (rows below the diagonal block updated left-looking inside POTRF's block columns) crossed with the
outer block width (knob 14, tiles): ms per headline
evaluation, POTRF alone (knob 4 = 0), objective / gradient drift against the right-looking order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpscore_b200 import api, synth
ctx = api.Context(0)
s = torch.cuda.Stream(); ctx.set_stream(s)
X, y = synth.kin40k_like(10000); theta = synth.hyper_point("P1")
ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
base = None
combos = [(1, 8, 0), (1, 8, 1), (1, 12, 0), (1, 12, 1), (1, 16, 0), (1, 16, 1), (1, 6, 1), (1, 10, 1), (0, 8, 0), (0, 8, 1), (0, 12, 1)]
for overlap, ob, left in combos:
    if True:
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 4, overlap))
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 14, ob))
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 13, left))
        for _ in range(2): v, g = ctx.full_eval(theta, "crps")
        e0.record(s)
        for _ in range(5): v, g = ctx.full_eval(theta, "crps")
        e1.record(s); s.synchronize()
        if base is None: base = (v, g.copy())
        st = ctx.last_stage_ms()
        print("overlap %d OB %2d left-looking %d: %.2f ms/eval  potrf %.2f trtri %.2f | obj drift %.2e grad drift %.2e" % (
            overlap, ob, left, e0.elapsed_time(e1) / 5, st["potrf"], st["trtri"], abs(v - base[0]) / abs(base[0]),
            np.max(np.abs(g - base[1])) / np.max(np.abs(base[1]))), flush=True)
