"""A/B of knob 15 (X phases of the inversion tree's right spine released row group by row group behind POTRF): ms per
headline evaluation and bit-level drift (the products are the same, only their release order changes)."""
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
for rw, pct in ((0, 50), (1, 50), (0, 50), (1, 50), (1, 60), (1, 40)):
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 9, pct))
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 15, rw))
    for _ in range(2): v, g = ctx.full_eval(theta, "crps")
    e0.record(s)
    for _ in range(5): v, g = ctx.full_eval(theta, "crps")
    e1.record(s); s.synchronize()
    if base is None: base = (v, g.copy())
    st = ctx.last_stage_ms()
    print("row-wise %d split %d%%: %.2f ms/eval  potrf+trtri %.2f | obj drift %.2e grad drift %.2e" % (
        rw, pct, e0.elapsed_time(e1) / 5, st["potrf"] + st["trtri"], abs(v - base[0]) / abs(base[0]),
        np.max(np.abs(g - base[1])) / np.max(np.abs(base[1]))), flush=True)
