"""How much low-precision tensor throughput stands behind the FP64-emulation route of DESIGN.md §8?  Library
GEMMs only (cuBLASLt through torch): INT8 -> INT32, BF16, and FP64 on the same shapes."""
import time

import torch


def bench(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


for n in (8192, 16384):
    a8 = torch.randint(-127, 127, (n, n), dtype=torch.int8, device="cuda")
    b8 = torch.randint(-127, 127, (n, n), dtype=torch.int8, device="cuda")
    try:
        t = bench(lambda: torch._int_mm(a8, b8))
        print("n=%5d int8 -> int32: %.1f TOP/s" % (n, 2 * n ** 3 / t / 1e12))
    except Exception as exc:  # noqa: BLE001
        print("int8 gemm not available:", exc)
    ab = torch.randn(n, n, dtype=torch.bfloat16, device="cuda")
    bb = torch.randn(n, n, dtype=torch.bfloat16, device="cuda")
    t = bench(lambda: ab @ bb)
    print("n=%5d bf16: %.1f TFLOP/s" % (n, 2 * n ** 3 / t / 1e12))
    if n == 8192:
        ad = torch.randn(n, n, dtype=torch.float64, device="cuda")
        t = bench(lambda: ad @ ad, reps=3)
        print("n=%5d fp64: %.1f TFLOP/s" % (n, 2 * n ** 3 / t / 1e12))
