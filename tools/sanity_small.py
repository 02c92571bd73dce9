"""Small end-to-end pass over every CUDA path (a quick sanity run on a GPU box): full GP (all objectives), FITC fused
and matrix form, prediction, metrics, grid."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpscore_b200 import api, synth  # noqa: E402

ctx = api.Context(0)
X, y = synth.kin40k_like(900, seed=1)
Xs, ys = synth.kin40k_like(300, seed=2)
theta = synth.hyper_point("P1")
ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
for s in ("crps", "logs", "nlml", "dss"):
    print("full", s, ctx.full_eval(theta, s)[0])
m, v = ctx.full_predict(theta, torch.from_numpy(Xs).cuda())
print("metrics", ctx.test_metrics(m, v, torch.from_numpy(ys).cuda(), torch.from_numpy(y).cuda())["crps"])
U = synth.inducing_init(20, seed=3)
for s in ("crps", "logs", "nlml", "dss", "kc"):
    print("fitc20", s, ctx.fitc_eval(theta, U, s)[0])
m, v = ctx.fitc_predict(theta, U, torch.from_numpy(Xs).cuda())
U2 = X[:150] + 0.01
for s in ("crps", "nlml"):
    print("fitc150", s, ctx.fitc_eval(theta, U2, s)[0])
m, v = ctx.fitc_predict(theta, U2, torch.from_numpy(Xs).cuda())
print("fitc150 sharded", ctx.fitc_eval_sharded(theta, U2, "crps", 900, lambda t: t)[0])
x1 = np.linspace(-3, 3, 60)
y1 = np.sin(x1)
print("grid", ctx.grid_eval(x1, y1, np.array([0.5, 1.0]), np.array([0.1, 0.3]), "crps"))
ctx.close()
print("done")
