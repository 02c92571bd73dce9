"""POTRF stage time at small N (the trailing updates are negligible there): time per 128-wide tile step
of the diagonal-block chain."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpscore_b200 import api, synth  # noqa: E402

ctx = api.Context(0)
ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 4, 0))
theta = synth.hyper_point("P1")
strip = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 7, strip))
print("chain strip policy", strip)
for N in (512, 2048):
    X, y = synth.kin40k_like(N)
    ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
    for _ in range(3):
        ctx.full_eval(theta, "crps")
    best = None
    for _ in range(5):
        ctx.full_eval(theta, "crps")
        st = ctx.last_stage_ms()
        best = st if best is None or st["potrf"] < best["potrf"] else best
    nb = (N + 127) // 128
    print("N=%5d tiles=%3d potrf %.3f ms = %.1f us/tile  trtri %.3f lauum %.3f symprod %.3f" % (
        N, nb, best["potrf"], best["potrf"] * 1e3 / nb, best["trtri"], best["lauum"], best["symprod"]))
