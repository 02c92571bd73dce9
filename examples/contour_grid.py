#!/usr/bin/env python
"""The grid sweep of contour-plot.R (CP:88-144) on the GPU: NLML, LOO-CRPS, in-sample ("wrong") CRPS
and LOO log score on the 50 x 50 (length scale, noise s.d.) grid; writes the four matrices in the
layout `contour(noise_range, l_range, matrix)` expects (CP:114-141) as .npy and .csv."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpscore_b200.api as gp  # noqa: E402
from gpscore_b200 import dist as gd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20)            # num_train, CP:33
    ap.add_argument("--grid", type=int, default=50)         # length.out, CP:88 / CP:109
    ap.add_argument("--out", default="contour_out")
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    x = np.linspace(-6, 6, args.n)                                                   # CP:35
    K = np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2) + 1e-10 * np.eye(args.n)       # rbf(l=1,k=1), CP:36
    y = np.linalg.cholesky(K) @ rng.standard_normal(args.n) + 0.1 * rng.standard_normal(args.n)   # CP:37-38
    l_range = np.linspace(0.01, 2, args.grid)                                        # CP:88
    noise_range = np.linspace(0.01, 1, args.grid)                                    # CP:109
    Lg, Sg = np.meshgrid(l_range, noise_range, indexing="ij")
    ctx = gp.default_context()
    os.makedirs(args.out, exist_ok=True)
    for which in ("nlml", "wrong_crps", "crps", "logs"):                             # CP:113, 121, 129, 138
        vals = ctx.grid_eval(x, y, Lg.ravel(), Sg.ravel(), which)
        mat = gd.grid_matrix(vals, args.grid, args.grid)                             # rows = noise s.d., cols = l
        np.save(os.path.join(args.out, "ma_%s.npy" % which), mat)
        np.savetxt(os.path.join(args.out, "ma_%s.csv" % which), mat, delimiter=",")
        i, j = np.unravel_index(np.argmin(mat), mat.shape)
        print("%-10s min %.6f at noise s.d. %.3f, length scale %.3f" % (which, mat[i, j], noise_range[i], l_range[j]))


if __name__ == "__main__":
    main()
