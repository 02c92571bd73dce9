"""One full-GP evaluation + one prediction chunk at N rows: the program under ncu for the HBM-class kernels
(gram_sym, grad_contract, symv, loo_score, scale_cols, predict_rows)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpscore_b200 import api, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
X, y, Xs, ys = synth.kin40k_like(N, 4096)
theta = synth.hyper_point("P1")
ctx = api.Context(0)
ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
for _ in range(2):
    print(ctx.full_eval(theta, "crps")[0])
m, v = ctx.full_predict(theta, torch.from_numpy(Xs).cuda())
print(float(m.sum()), float(v.sum()))
