"""Large-n grid sweep (CP:109-144 shape at BASELINE configs[4] size): time per grid point."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpscore_b200 import api  # noqa: E402

ctx = api.Context(0)
rng = np.random.default_rng(0)
for n in (512, 2048, 4096):
    x = np.sort(rng.uniform(-5, 5, n))
    y = np.sin(x) + 0.3 * rng.standard_normal(n)
    G = 64
    ls = np.repeat(np.linspace(0.3, 2.0, 8), 8)
    sd = np.tile(np.linspace(0.1, 1.0, 8), 8)
    for which in ("crps", "nlml"):
        ctx.grid_eval(x, y, ls[:8], sd[:8], which)
        t0 = time.perf_counter()
        v = ctx.grid_eval(x, y, ls, sd, which)
        dt = time.perf_counter() - t0
        print("n=%5d %-5s %d points: %.1f ms = %.3f ms/point  (first %.6g last %.6g)" % (n, which, G, dt * 1e3, dt * 1e3 / G, v[0], v[-1]))
# the reference's own grid: 50 x 50 points at n <= 128 (one CTA per point, one launch)
n = 100
x = np.sort(rng.uniform(-5, 5, n))
y = np.sin(x) + 0.3 * rng.standard_normal(n)
ls = np.repeat(np.linspace(0.05, 3.0, 50), 50)
sd = np.tile(np.linspace(0.05, 2.0, 50), 50)
for which in ("nlml", "crps", "wrong_crps", "logs"):
    ctx.grid_eval(x, y, ls, sd, which)
    t0 = time.perf_counter()
    v = ctx.grid_eval(x, y, ls, sd, which)
    dt = time.perf_counter() - t0
    print("n=%5d %-10s 2500 points: %.2f ms" % (n, which, dt * 1e3))
