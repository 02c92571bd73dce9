"""A/B of the TRTRI node split (knob 9) on the headline evaluation: ms per evaluation (CUDA events), stage times,
objective drift against halving."""
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
for pct in [int(a) for a in sys.argv[1:]] or [50, 60, 66, 72, 78, 84]:
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 9, pct))
    for _ in range(2): v, g = ctx.full_eval(theta, "crps")
    e0.record(s)
    for _ in range(5): v, g = ctx.full_eval(theta, "crps")
    e1.record(s); s.synchronize()
    if base is None: base = (v, g.copy())
    st = ctx.last_stage_ms()
    print("split %d%%: %.2f ms/eval  potrf+trtri %.2f  lauum %.2f  symprod %.2f | obj drift %.2e grad drift %.2e" % (
        pct, e0.elapsed_time(e1) / 5, st["potrf"] + st["trtri"], st["lauum"], st["symprod"], abs(v - base[0]) / abs(base[0]),
        np.max(np.abs(g - base[1])) / np.max(np.abs(base[1]))), flush=True)
