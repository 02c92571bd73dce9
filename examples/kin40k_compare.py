#!/usr/bin/env python
"""The reference experiment, ported onto the fused CUDA objectives.

Mirrors kin40k-FULL-compare.py (KF:190-299, 312-399, 405-483) and
KIN40K-COMPARE-ALL-FITC-20.py (K20:184-304, 315-412, 417-518): per trial, fit the
hyper-parameters by fixed-step gradient descent on LOO-CRPS, NLML and the LOO log score
(same learning rates and iteration counts as the scripts unless --itr-scale shrinks them),
predict on the test set and tabulate mse / SMSE / logs / CRPS / MSLL / +-2 sd coverage.
The KIN40K workbook is not part of the reference repository (KF:141), so the data is the seeded
synthetic stand-in of gpscore_b200.synth.  Needs a CUDA device.

    python examples/kin40k_compare.py --model full --trials 2 --itr-scale 0.1
    python examples/kin40k_compare.py --model fitc --trials 2 --itr-scale 0.05
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpscore_b200.api as gp  # noqa: E402
from gpscore_b200 import synth  # noqa: E402

# (score, iterations, learning rate, learning rate of the inducing inputs) as in the scripts
FULL_RUNS = [("crps", 400, 1.0, None), ("nlml", 400, 0.0005, None), ("logs", 500, 0.05, None),         # KF:220,238 / 312,328 / 405,415
             ("dss", 150, 0.001, None)]                                                                   # KF:487,498 (4-fold DSS)
FITC_RUNS = [("crps", 2000, 1.0, 1.0), ("nlml", 3000, 0.0001, 0.001), ("logs", 3000, 0.2, 0.2),         # K20:207,220 / 315,326 / 417,430
             ("dss", 3000, 0.001, 0.001), ("kc", 3000, 0.1, 0.1)]                                         # K20:523,537 / 655,668


def fit(train_x, train_y, score, itr, lr, lr2, fitc, m, seed):
    torch.manual_seed(seed)
    d = train_x.shape[1]
    para_l = torch.rand(1, d, dtype=torch.float64, requires_grad=True)          # KF:226
    para_k = torch.tensor([1.0], dtype=torch.float64, requires_grad=True)       # KF:323
    para_noise = torch.tensor([1.0], dtype=torch.float64, requires_grad=True)
    inducing_x = torch.rand(m, d, dtype=torch.float64, requires_grad=True) if fitc else None   # K20:215
    for i in range(itr):
        if fitc:
            obj = gp.fitc_loo_objective(train_x, train_y, inducing_x, para_k, para_l, para_noise, score)
        else:
            obj = gp.full_loo_objective(train_x, train_y, para_k, para_l, para_noise, score)
        obj.backward()
        with torch.no_grad():                                                    # KF:254-260 / K20:243-251
            para_l -= lr * para_l.grad
            para_k -= lr * para_k.grad
            para_noise -= lr * para_noise.grad
            para_l.grad.zero_()
            para_k.grad.zero_()
            para_noise.grad.zero_()
            if fitc:
                inducing_x -= lr2 * inducing_x.grad
                inducing_x.grad.zero_()
    return para_k, para_l, para_noise, inducing_x, float(obj.detach().reshape(-1)[0])


def fit_on_device(ctx, train_x, train_y, score, itr, lr, lr2, fitc, m, seed):
    """The same loop in ONE library call: theta and the inducing inputs never leave the device between steps
    (gps_fitc_descend / gps_full_descend).  Same initialisation as `fit`."""
    torch.manual_seed(seed)
    d = train_x.shape[1]
    para_l = torch.rand(1, d, dtype=torch.float64)
    theta = np.concatenate([[1.0], para_l.numpy().ravel(), [1.0]])
    ctx.set_data(train_x, train_y)
    if fitc:
        U0 = torch.rand(m, d, dtype=torch.float64).numpy()
        theta, U, trace = ctx.fitc_descend(theta, U0, score, lr, lr2, itr)
    else:
        theta, trace = ctx.full_descend(theta, score, lr, itr)
        U = None
    t = torch.from_numpy
    return (t(theta[:1]), t(theta[1:-1]).reshape(1, -1), t(theta[-1:]), None if U is None else t(U), float(trace[-1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", choices=["full", "fitc"], default="full")
    ap.add_argument("--trials", type=int, default=2)          # TT = 30 / 10 in the scripts (KF:149, K20:145)
    ap.add_argument("--n-train", type=int, default=500)       # KF:196-213
    ap.add_argument("--n-test", type=int, default=500)
    ap.add_argument("--m", type=int, default=20)              # K20:205
    ap.add_argument("--itr-scale", type=float, default=0.1, help="fraction of the scripts' iteration counts")
    ap.add_argument("--on-device", action="store_true", help="run each optimiser loop as one library call (device-resident)")
    args = ap.parse_args()
    fitc = args.model == "fitc"
    runs = FITC_RUNS if fitc else FULL_RUNS
    if fitc and args.m > 32:   # beyond 32 inducing points the library runs its GEMM formulation: crps / nlml / logs only
        runs = [r for r in runs if r[0] not in ("dss", "kc")]
    table = {r[0]: [] for r in runs}
    for j in range(args.trials):
        X, y, Xs, ys = synth.kin40k_like(args.n_train, args.n_test, seed=100 * j)   # random.seed(j*100), KF:194
        train_x, train_y = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
        test_x, test_y = torch.from_numpy(Xs).cuda(), torch.from_numpy(ys).cuda()
        for score, itr, lr, lr2 in runs:
            itr = max(1, int(itr * args.itr_scale))
            if args.on_device and score in ("crps", "logs", "nlml"):
                pk, pl, pn, U, last = fit_on_device(gp.default_context(), train_x, train_y, score, itr, lr, lr2, fitc, args.m, 100 * j)
            else:
                pk, pl, pn, U, last = fit(train_x, train_y, score, itr, lr, lr2, fitc, args.m, 100 * j)
            mean, var = gp.predict_diag(train_x, train_y, test_x, pk, pl, pn, inducing_x=U)     # KF:267-273
            met = gp.test_metrics(mean, var, test_y, train_y)                                    # KF:276-292
            table[score].append([met[k] for k in ("mse", "smse", "logs", "crps", "msll", "coverage")])
            print("trial %d %-4s itr %4d objective %.6f  " % (j, score, itr, last) +
                  "  ".join("%s %.4f" % kv for kv in met.items()))
    print("\nmeans over trials (KF:739-776):")
    for score, rows in table.items():
        m = np.mean(np.array(rows), axis=0)
        print("  fitted by %-4s  mse %.4f  smse %.4f  logs %.4f  crps %.4f  msll %.4f  coverage %.4f" % (score, *m))


if __name__ == "__main__":
    main()
