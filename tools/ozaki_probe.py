"""Gate check for the FP64-emulation route of DESIGN.md §8 (round-1 verdict item 8), at LIBRARY level only:
an Ozaki-style split of an fp64 GEMM C = A B' into S x S int8 slice products (torch._int_mm = cuBLASLt IMMA on the
tcgen05 tensor cores), exact int32 accumulation, fp64 recombination.  It answers two questions before any kernel is
written: (1) what error does the scheme leave against a native DGEMM on matrices shaped like this path's operands
(K^-1 of a low-noise Gram and K^-1 diag(dbar)), (2) how many slice products does 1e-13 need, and what FP64-equivalent
rate do they run at when only the int8 GEMM time is counted (an upper bound: slicing and recombination passes are
extra, a fused kernel would hide most of them).

Not on the product path.  Writes gpurun_out/ozaki_probe.json."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

BETA = 6   # bits per slice: k * 2^(2 beta) < 2^31 for k = 8192 leaves room to add 8 same-scale products in int32


def split(M, S, dim):
    """M ~= sum_s 2^(e - s*BETA) * Q_s with integer Q_s in [-2^BETA, 2^BETA], e per row (dim=1) or per column (dim=0)."""
    amax = M.abs().amax(dim=dim, keepdim=True).clamp_min(1e-300)
    e = torch.ceil(torch.log2(amax))
    R = M / torch.exp2(e)                    # |R| <= 1
    Q = []
    for s in range(1, S + 1):
        q = torch.round(R * (2.0 ** BETA))   # |q| <= 2^BETA
        R = R * (2.0 ** BETA) - q            # |R| <= 1/2
        Q.append(q.to(torch.int8))
    return e, Q


def emulated_gemm(A, B, S, count_only=False):
    """C = A @ B.T through int8 slices; products with s + t > S + 1 are dropped (below the target precision)."""
    ea, QA = split(A, S, 1)
    eb, QB = split(B, S, 1)
    n, m = A.shape[0], B.shape[0]
    C = torch.zeros(n, m, dtype=torch.float64, device=A.device)
    nprod = 0
    for lvl in range(2, S + 2):              # lvl = s + t: all products of a level share the scale 2^(-lvl * BETA)
        acc = None
        for s in range(1, S + 1):
            t = lvl - s
            if t < 1 or t > S:
                continue
            nprod += 1
            if count_only:
                continue
            p = torch._int_mm(QA[s - 1], QB[t - 1].t())   # int32, exact (B slice passed column-major)
            acc = p if acc is None else acc + p
        if not count_only and acc is not None:
            C += acc.to(torch.float64) * (2.0 ** (-lvl * BETA))
    if count_only:
        return nprod
    return C * torch.exp2(ea) * torch.exp2(eb).t(), nprod


def time_int8(n, reps=5):
    a = torch.randint(-64, 64, (n, n), dtype=torch.int8, device="cuda")
    b = torch.randint(-64, 64, (n, n), dtype=torch.int8, device="cuda")
    torch._int_mm(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch._int_mm(a, b)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    from gpscore_b200 import synth
    out = {"beta_bits": BETA}
    n = 4096
    X, y = synth.kin40k_like(n)
    th = synth.hyper_point("P2")             # low-noise point: cond(K) large, K^-1 badly scaled
    Xd = torch.from_numpy(X / np.exp(th[1:-1])).cuda()
    r2 = torch.cdist(Xd, Xd) ** 2
    K = np.exp(th[0]) * torch.exp(-0.5 * r2) + np.exp(th[-1]) * torch.eye(n, dtype=torch.float64, device="cuda")
    Kinv = torch.linalg.inv(K)
    dbar = torch.randn(n, dtype=torch.float64, device="cuda")
    cases = {"random N(0,1)": (torch.randn(n, n, dtype=torch.float64, device="cuda"), torch.randn(n, n, dtype=torch.float64, device="cuda")),
             "K^-1 diag(dbar) x K^-1 (P2, N=4096)": (Kinv * dbar, Kinv)}
    for name, (A, B) in cases.items():
        ref = A @ B.t()
        scale = (A.abs() @ B.abs().t()).amax()
        rows = {}
        for S in (5, 6, 7, 8, 9):
            C, nprod = emulated_gemm(A, B, S)
            rows["S=%d" % S] = {"slice_products": nprod, "err_vs_absA_absB": float((C - ref).abs().max() / scale),
                                "err_vs_max_abs_C": float((C - ref).abs().max() / ref.abs().max())}
        # what a native DGEMM itself leaves against a higher-precision reference is ~1e-16 |A||B| sqrt(k)
        out[name] = rows
    t8 = time_int8(8192)
    tops = 2 * 8192 ** 3 / t8 / 1e12
    out["int8_gemm_8192_TOPs"] = tops
    eq = {}
    for S in (6, 7, 8, 9):
        nprod = sum(1 for lvl in range(2, S + 2) for s in range(1, S + 1) if 1 <= lvl - s <= S)
        eq["S=%d" % S] = {"slice_products": nprod, "TFLOPs_equiv": tops / nprod}
    out["fp64_equivalent_TFLOPs_gemm_time_only"] = eq
    ad = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    ad @ ad
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ad @ ad; e1.record(); torch.cuda.synchronize()
    out["cublas_dgemm_8192_TFLOPs"] = 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "ozaki_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
