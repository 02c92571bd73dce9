"""Bit-level A/B of the tile-GEMM policies on the full-GP evaluation (TMA ring vs cp.async ring must agree exactly)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpscore_b200 import api, synth
ctx = api.Context(0)
for n in (1000, 4000, 10000):
    X, y = synth.kin40k_like(n); theta = synth.hyper_point("P1")
    ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
    res = {}
    for v in (6, 9, 9, 10, 10, 11, 11):
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 0, v))
        val, g = ctx.full_eval(theta, "crps")
        res.setdefault(v, []).append((val, g.copy()))
    base = res[6][0]
    for v in (9, 10, 11):
        for k, (val, g) in enumerate(res[v]):
            print("N=%d variant %d run %d: obj diff %.3e grad diff %.3e" % (n, v, k, abs(val - base[0]), np.max(np.abs(g - base[1]))), flush=True)
# raw GEMM equality per kind on random data, several sizes
for n in (512, 2048):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda"); B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for kind in (0, 1, 2):
        outs = []
        for v in (6, 9, 10, 11):
            ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 0, v))
            Cd = torch.zeros(n, n, dtype=torch.float64, device="cuda")
            ctx._check(ctx._lib.gps_dbg_gemm(ctx._h, kind, A.data_ptr(), B.data_ptr(), Cd.data_ptr(), n, n, n, 1.0, 0.0, None, 0))
            outs.append(Cd.clone())
        print("gemm n=%d kind %d: max |v - v6| for v = 9, 10, 11: %s" % (n, kind, ["%.3e" % float((o - outs[0]).abs().max()) for o in outs[1:]]), flush=True)
