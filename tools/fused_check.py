"""Quick diagnostic of the fused FITC path on one GPU: errors against the CPU Woodbury oracle and timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpscore_b200 import api, synth
from oracle import gp_oracle as O, woodbury as W

def rel(a, b):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

ctx = api.Context(0)
for n, m in ((300, 20), (1000, 5), (4097, 31), (10000, 20)):
    X, y = synth.kin40k_like(n, seed=n)
    th = synth.hyper_point("P1"); U = synth.inducing_init(m)
    ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
    for sc in ("crps", "logs", "nlml"):
        try:
            v, g, gu = ctx.fitc_eval(th, U, sc)
            ov, og, ogu = W.fitc_obj_grad(X, y, U, th, O.SCORES[sc])[:3]
            print("N=%d M=%d %s obj %.3e grad %.3e gradU %.3e" % (n, m, sc, abs(v - ov) / abs(ov), rel(g, og), rel(gu, ogu)), flush=True)
            if rel(g, og) > 1e-6:
                print("   g ", g, "\n   og", og)
        except Exception as e:
            print("N=%d M=%d %s EXC %r" % (n, m, sc, e), flush=True)
s = torch.cuda.Stream(); ctx.set_stream(s)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in (10000, 100000, 1000000):
    X, y = synth.kin40k_like(n, seed=7)
    th = synth.hyper_point("P1"); U = synth.inducing_init(20)
    ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
    for variant in (1, 2):
        ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 2, variant))
        for _ in range(5): ctx.fitc_eval(th, U, "crps")
        reps = 200 if n <= 100000 else 30
        t0 = time.perf_counter(); e0.record(s)
        for _ in range(reps): ctx.fitc_eval(th, U, "crps")
        e1.record(s); s.synchronize(); t1 = time.perf_counter()
        print("N=%d variant %d: %.1f us/eval (events) %.1f us/eval (wall)" % (n, variant, e0.elapsed_time(e1) * 1e3 / reps, (t1 - t0) * 1e6 / reps), flush=True)
    ctx._check(ctx._lib.gps_dbg_set_variant(ctx._h, 2, 2))
    if n == 10000:
        t0 = time.perf_counter(); ctx.fitc_descend(th, U, "crps", 1e-3, 1e-3, 500); t1 = time.perf_counter()
        print("descend 500 iters: %.1f us/iter" % ((t1 - t0) * 1e6 / 500), flush=True)
        print("launch floor (3 launches + sync): %.1f us" % ctx.launch_floor_us(3, 100), flush=True)
# phase timeline (ns) of one evaluation at N = 1e4 and N = 1e6
import ctypes as C
for n in (10000, 1000000):
    X, y = synth.kin40k_like(n, seed=7)
    ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
    buf = (C.c_int64 * 48)()
    ctx.fitc_eval(th, U, "crps")
    ctx._check(ctx._lib.gps_dbg_fused_phases(ctx._h, buf))
    for _ in range(3): ctx.fitc_eval(th, U, "crps")
    ctx._check(ctx._lib.gps_dbg_fused_phases(ctx._h, buf))
    v = list(buf); t0 = v[0]
    names = ["start", "preamble", "rows", "cta-partial", "reduced", "finish"]
    for k in range(3):
        print("N=%d pass %d:" % (n, k + 1), "  ".join("%s %+.1f" % (names[j], (v[16 * k + j] - t0) / 1e3) for j in range(6) if v[16 * k + j]), flush=True)
    print("   preamble detail p1: params+Us %+.1f  Kuu %+.1f  chol %+.1f  inverse %+.1f | p2: loads %+.1f chol %+.1f inverse %+.1f | p3: loads+SW %+.1f  LCbar %+.1f  adjoint %+.1f" % (
        tuple((v[j] - v[0]) / 1e3 for j in (5, 6, 7, 8)) + tuple((v[j] - v[16]) / 1e3 for j in (21, 22, 23)) + tuple((v[j] - v[32]) / 1e3 for j in (44, 45, 46))))
    print("   finish detail: loaded %+.1f  S assembled %+.1f  adjoint done %+.1f  theta grads done %+.1f" % tuple((v[40 + j] - t0) / 1e3 for j in range(4)))
