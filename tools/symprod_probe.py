"""Why does the symmetric product run below the plain tile GEMM?  Same policy, n = 10112 (79 tiles):
plain NT GEMM, NT with the k-scaling vector, A = B operand vs distinct operands."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpscore_b200 import api  # noqa: E402

ctx = api.Context(0)
for n in (8192, 10112):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    dv = torch.randn(n, dtype=torch.float64, device="cuda")
    for name, a, b, d in (("A*A' ", A, A, None), ("A*B' ", A, B, None), ("A*D*A'", A, A, dv), ("A*D*B'", A, B, dv)):
        best = 1e9
        for _ in range(3):
            ctx._check(ctx._lib.gps_dbg_gemm(ctx._h, 0, a.data_ptr(), b.data_ptr(), C.data_ptr(), n, n, n, 1.0, 0.0,
                                             d.data_ptr() if d is not None else None, 0))
            best = min(best, ctx.last_gemm_ms()[0])
        print("n=%5d %s full square: %.2f ms  %.2f TFLOP/s" % (n, name, best, 2 * n ** 3 / (best * 1e-3) / 1e12))
