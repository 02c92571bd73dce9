"""One matrix-form FITC evaluation (for ncu launch lists): N, M from argv."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpscore_b200 import api, synth  # noqa: E402

N, M = int(sys.argv[1]), int(sys.argv[2])
score = sys.argv[3] if len(sys.argv) > 3 else "crps"
ctx = api.Context(0)
X, y = synth.kin40k_like(N, seed=7)
ctx.set_data(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
rng = np.random.default_rng(1)
U = X[rng.choice(N, M, replace=False)] + 0.01 * rng.standard_normal((M, X.shape[1]))
print(ctx.fitc_eval(synth.hyper_point("P1"), U, score)[0])
