"""Timeline of the factorisation lanes (gps_dbg_trace) of one headline evaluation: when each outer step's
diagonal-block chain, below-rows panel, trailing update and inversion merges finish."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpscore_b200 import api, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
X, y = synth.kin40k_like(N)
theta = synth.hyper_point("P1")
ctx = api.Context(0)
ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 4, mode))
ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 6, 1))
for _ in range(3):
    ctx.full_eval(theta, "crps")
codes = (C.c_int * 1024)()
ms = (C.c_double * 1024)()
n = ctx._lib.gps_dbg_trace(ctx._h, 1024, codes, ms)
lanes = {1: "chain", 2: "below", 3: "trail", 4: "merges"}
tab = {}
tiles = {}
for i in range(n):
    lane, step = codes[i] // 1000, codes[i] % 1000
    if lane in lanes:
        tab.setdefault(step, {})[lanes[lane]] = ms[i]
    elif lane in (5, 6):
        tiles.setdefault(step, {})["chain" if lane == 5 else "below"] = ms[i]
    elif lane == 7:
        tab.setdefault(step, {})["trailA"] = ms[i]
print("step   chain   below   trail  merges   (ms since start; mode %d)" % mode)
for step in sorted(tab):
    r = tab[step]
    print("%4d %7.2f %7.2f %7.2f %7s" % (step, r.get("chain", float("nan")), r.get("below", float("nan")),
                                        r.get("trail", float("nan")), ("%.2f" % r["merges"]) if "merges" in r else "-"))
if len(sys.argv) > 3:
    print("tile   chain   below   (ms since start; 128-wide steps of the last outer blocks)")
    for k in sorted(tiles)[-24:]:
        print("%4d %7.3f %7.3f" % (k, tiles[k].get("chain", float("nan")), tiles[k].get("below", float("nan"))))
    print("trailA done:", {o: round(tab[o]["trailA"], 3) for o in sorted(tab)[-6:] if "trailA" in tab[o]})
print("stages", {k: round(v, 2) for k, v in ctx.last_stage_ms().items()})
