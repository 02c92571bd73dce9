"""FITC scaling sweep (BASELINE.json configs[4]): synthetic 8-D, N rows, M = 20 ... 1024 inducing points.
M <= 32 runs the fused row kernels, M > 32 the matrix form (csrc/gps_fitc_large.cu).  Reports ms per
objective+gradient evaluation and the useful FP64 rate of the matrix form (10 M^2 N flops: eight big
products with their triangular / symmetric halves not counted)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpscore_b200 import api, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
Ms = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [20, 64, 128, 256, 512, 1024]
score = sys.argv[3] if len(sys.argv) > 3 else "crps"
ctx = api.Context(0)
X, y = synth.kin40k_like(N, seed=7)
ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
theta = synth.hyper_point("P1")
rng = np.random.default_rng(1)
out = []
for M in Ms:
    U = X[rng.choice(N, M, replace=False)] + 0.01 * rng.standard_normal((M, X.shape[1]))
    for _ in range(2):
        v = ctx.fitc_eval(theta, U, score)[0]
    torch.cuda.synchronize()
    reps = 20 if M <= 64 else (5 if M <= 256 else 3)
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.fitc_eval(theta, U, score)
    dt = (time.perf_counter() - t0) / reps
    launches = (ctx.launch_count() - l0) // reps
    Mp = (M + 127) // 128 * 128
    row = {"N": N, "M": M, "score": score, "ms_per_eval": dt * 1e3, "launches": launches, "obj": v}
    if M > 32:
        row["useful_tflops"] = 10.0 * M * M * N / dt / 1e12
        row["padded_tflops"] = 10.0 * Mp * Mp * N / dt / 1e12
    out.append(row)
    print(json.dumps(row), flush=True)
