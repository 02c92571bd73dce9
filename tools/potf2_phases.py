"""Phase timing (clock64) of the diagonal-block kernel: one factorisation of a 128 x 128 SPD block."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpscore_b200 import api  # noqa: E402

ctx = api.Context(0)
buf = (C.c_int64 * 17)()
ctx._check(ctx._lib.gps_dbg_potf2_phases(ctx._h, buf))   # arm
rng = np.random.default_rng(0)
G = rng.standard_normal((128, 160))
A = torch.from_numpy(G @ G.T / 128 + 0.5 * np.eye(128)).cuda()
L = torch.empty_like(A)
for _ in range(3):
    ctx._check(ctx._lib.gps_dbg_factor(ctx._h, A.data_ptr(), 128, L.data_ptr(), None, None))
ctx._check(ctx._lib.gps_dbg_potf2_phases(ctx._h, buf))
c = list(buf)
names = ["load"] + [n % k for k in range(4) for n in ("chol32[%d]", "panel[%d]", "update[%d]")] + ["inv32", "inv offdiag", "writeback"]
prev = 0
for n, v in zip(names, c[1:]):
    print("%-14s %8d cycles" % (n, v - prev))
    prev = v
print("total %d cycles = %.1f us at 1.965 GHz" % (c[16], c[16] / 1965.0))
