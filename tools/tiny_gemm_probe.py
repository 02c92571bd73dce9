"""Cost of the few-tile, k = 128 GEMM launches on POTRF's serial chain: 1-, 3-, 6-, 76- and 228-tile products
(CUDA events around each launch), next to the per-launch cost of a dependent chain of tiny torch kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpscore_b200 import api  # noqa: E402

ctx = api.Context(0)
st = torch.cuda.Stream()
ctx.set_stream(st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(st):
    z = torch.zeros(8, device="cuda")
    for _ in range(50):
        z.add_(1.0)
    e0.record(st)
    for _ in range(1000):
        z.add_(1.0)
    e1.record(st)
    st.synchronize()
    print("torch tiny kernel: %.2f us per dependent launch" % (e0.elapsed_time(e1)))
ctx.set_gemm_timing(True)
for mt, nt in ((1, 1), (3, 1), (3, 2), (76, 1), (76, 3)):
    A = torch.randn(mt * 128, 128, dtype=torch.float64, device="cuda")
    B = torch.randn(nt * 128, 128, dtype=torch.float64, device="cuda")
    Cm = torch.zeros(mt * 128, nt * 128, dtype=torch.float64, device="cuda")
    for beta in (0.0, 1.0):
        best = 1e9
        for _ in range(30):
            ctx._check(ctx._lib.gps_dbg_gemm(ctx._h, 0, A.data_ptr(), B.data_ptr(), Cm.data_ptr(), mt * 128, nt * 128, 128,
                                             1.0, beta, None, 0))
            best = min(best, ctx.last_gemm_ms()[0])
        print("gemm %2d x %d tiles, k=128, beta=%g: %.2f us (events around the launch, best of 30)" % (mt, nt, beta, best * 1e3))
