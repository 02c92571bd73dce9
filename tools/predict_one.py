"""One full-GP prediction call (for ncu launch lists): N = 10000 training rows, T = 30000 test rows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpscore_b200 import api, synth
ctx = api.Context(0)
X, y, Xs, ys = synth.kin40k_like(10000, 30000)
ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
m, v = ctx.full_predict(synth.hyper_point("P1"), torch.from_numpy(Xs).cuda())
print(float(m[0]), float(v[0]))
