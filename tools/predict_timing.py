"""kin40k-FULL prediction (KF:267-273): ms for N = 10000 training rows, T = 30000 test rows, one B200."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpscore_b200 import api, synth
ctx = api.Context(0)
stream = torch.cuda.Stream(); ctx.set_stream(stream)
theta = synth.hyper_point("P1")
X, y, Xs, ys = synth.kin40k_like(10000, 30000)
ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
Xsd = torch.from_numpy(Xs).cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
m, v = ctx.full_predict(theta, Xsd)
e0.record(stream)
for _ in range(3): m, v = ctx.full_predict(theta, Xsd)
e1.record(stream); torch.cuda.synchronize()
print("full_predict N=10000 T=30000: %.1f ms   mean[0] %.12g var[0] %.12g" % (e0.elapsed_time(e1) / 3, float(m[0]), float(v[0])))
