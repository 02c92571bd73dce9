"""Measure the FP64 roofline denominators on this B200: cuBLAS DGEMM 8192^3 (burst and sustained),
raw DMMA / DFMA issue rates, and time one full-GP evaluation with a stage breakdown.
Writes gpurun_out/fp64_peaks.json.  Run on the GPU box only."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpscore_b200 import api, synth  # noqa: E402

out = {}
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    torch.matmul(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
out["cublas_dgemm_8192_burst_tflops"] = 2 * n ** 3 / (best * 1e-3) / 1e12
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 0
e0.record()
t0 = time.time()
while time.time() - t0 < 3.0:
    for _ in range(5):
        torch.matmul(a, b)
    reps += 5
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
out["cublas_dgemm_8192_sustained_tflops"] = reps * 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
del a, b

ctx = api.Context(0)
x, y = C.c_double(), C.c_double()
ctx._check(ctx._lib.gps_dbg_fp64_peak(ctx._h, 20000, C.byref(x), C.byref(y)))
out["dmma_issue_tflops"], out["dfma_issue_tflops"] = x.value, y.value

# own tile GEMM, 8192^3 NT
n = 8192
A = torch.randn(n, n, dtype=torch.float64, device="cuda")
Cd = torch.zeros(n, n, dtype=torch.float64, device="cuda")
for kind in (0, 1, 2):
    best = 1e9
    for _ in range(4):
        ctx._check(ctx._lib.gps_dbg_gemm(ctx._h, kind, A.data_ptr(), A.data_ptr(), Cd.data_ptr(), n, n, n, 1.0, 0.0, None, 0))
        best = min(best, ctx.last_gemm_ms()[0])
    out["own_gemm_kind%d_8192_tflops" % kind] = 2 * n ** 3 / (best * 1e-3) / 1e12
del A, Cd

for N in (2000, 10000):
    X, yv = synth.kin40k_like(N)
    theta = synth.hyper_point("P1")
    ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(yv).cuda())
    for score in ("crps", "nlml"):
        ctx.full_eval(theta, score)
        ts = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ctx.full_eval(theta, score)
            ts.append(time.perf_counter() - t0)
        gm, gl = ctx.last_gemm_ms()
        out["full_eval_N%d_%s_ms" % (N, score)] = min(ts) * 1e3
        out["full_eval_N%d_%s_gemm_ms" % (N, score)] = gm
        out["full_eval_N%d_%s_gemm_launches" % (N, score)] = gl
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fp64_peaks.json"), "w"), indent=1)
