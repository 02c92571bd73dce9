"""A/B the tile-GEMM policies (gps_dbg_set_variant) on the GPU box: correctness vs torch.matmul,
8192^3 throughput per operand layout, and one full-GP evaluation at N = 10000."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpscore_b200 import api, synth  # noqa: E402

ctx = api.Context(0)
out = {}
n = 8192
A = torch.randn(n, n, dtype=torch.float64, device="cuda")
Cd = torch.zeros(n, n, dtype=torch.float64, device="cuda")
ref = None
X, y = synth.kin40k_like(10000)
theta = synth.hyper_point("P1")
variants = [int(v) for v in sys.argv[1:]] or [5, 6, 7, 8]
for v in variants:
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 0, v))
    for kind in (0, 1, 2):
        best = 1e9
        for _ in range(3):
            ctx._check(ctx._lib.gps_dbg_gemm(ctx._h, kind, A.data_ptr(), A.data_ptr(), Cd.data_ptr(), n, n, n, 1.0, 0.0, None, 0))
            best = min(best, ctx.last_gemm_ms()[0])
        out["v%d_kind%d_tflops" % (v, kind)] = 2 * n ** 3 / (best * 1e-3) / 1e12
        if kind == 0:
            if ref is None:
                ref = A[:512] @ A[:512].T
            err = float((Cd[:512, :512] - ref).abs().max() / ref.abs().max())
            out["v%d_relerr" % v] = err
    ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
    ctx.full_eval(theta, "crps")
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        val, g = ctx.full_eval(theta, "crps")
        ts.append(time.perf_counter() - t0)
    out["v%d_full_eval_ms" % v] = min(ts) * 1e3
    out["v%d_obj" % v] = val
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.full_eval(theta, "nlml")
        ts.append(time.perf_counter() - t0)
    out["v%d_nlml_eval_ms" % v] = min(ts) * 1e3
# diagonal-block kernel A/B at the default GEMM policy
for pv in (0, 1):
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 0, variants[0]))
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 1, pv))
    ctx.full_eval(theta, "crps")
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        val, g = ctx.full_eval(theta, "crps")
        ts.append(time.perf_counter() - t0)
    out["potf2v%d_full_eval_ms" % pv] = min(ts) * 1e3
    out["potf2v%d_obj" % pv] = val
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gemm_variants.json"), "w"), indent=1)
