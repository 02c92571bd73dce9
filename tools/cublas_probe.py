"""Run one cuBLAS DGEMM (torch.matmul fp64 8192^3) — under ncu this shows which kernel/tile shape
cuBLAS picks on B200 and its DMMA pipe utilisation, the comparator for gemm_tile_kernel."""
import torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
print("done", float(c[0, 0]))
