#!/usr/bin/env python
"""Wall-clock latency of the public calls at small batch sizes (host arrays in, lnL out):
    python dev/latency.py
B = 1 is the call an unmodified sampler makes (xpsi.Likelihood.__call__, xpsi/Likelihood.py:450)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from xpsi_b200 import synthetic as syn  # noqa: E402
from xpsi_b200.likelihood import Likelihood  # noqa: E402

w = bench.load_workload()
pipe = bench.make_pipeline(w, 512)
P = syn.m2_bench_thetas(0, 4096)
like = Likelihood(pipe, lambda pl, X: syn.m2_spot_batch(pl, X))
out = {}
for B in (1, 8, 64, 512):
    ts = []
    for r in range(24):
        blk = P[(r * B) % 3584:(r * B) % 3584 + B]
        t0 = time.perf_counter()
        spots = syn.m2_spot_batch(pipe, blk)
        t1 = time.perf_counter()
        lnL, st = pipe.eval_spots(spots)
        t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1))
    ts = np.array(ts[4:])
    out[B] = (np.median(ts[:, 0]) * 1e3, np.median(ts[:, 1]) * 1e3)
    print("B=%4d  host fill %.3f ms  eval_spots (H2D + kernels + D2H + sync) %.3f ms  -> %.0f evals/s"
          % (B, out[B][0], out[B][1], B / (out[B][0] + out[B][1]) * 1e3))
ts = []
for r in range(24):
    t0 = time.perf_counter()
    v = like(P[100 + r])
    ts.append(time.perf_counter() - t0)
print("Likelihood.__call__(p)  %.3f ms per call" % (np.median(ts[4:]) * 1e3))
