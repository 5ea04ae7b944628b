#!/usr/bin/env python
"""Wall time of the pieces of sampling.sweep on one GPU (fill, upload, run, download), after a small warm-up sweep
as in bench.py.    python dev/sweep_pieces.py [--n 100000]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--batch", type=int, default=512)
a = ap.parse_args()
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from xpsi_b200 import sampling, synthetic as syn  # noqa: E402
from xpsi_b200.likelihood import Likelihood  # noqa: E402

w = bench.load_workload()
pipe = bench.make_pipeline(w, a.batch)
like = Likelihood(pipe, lambda pl, X: syn.m2_spot_batch(pl, X))
P = syn.m2_bench_thetas(0, a.n)
sampling.sweep(like, P[:2 * a.batch], device=torch.device("cuda:0"))
for rep in range(2):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    spots = syn.m2_spot_batch(pipe, P); t.append(time.perf_counter())
    pipe.sweep_upload(spots); t.append(time.perf_counter())
    pipe.sweep_run(); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    lnL, st = pipe.sweep_download(); t.append(time.perf_counter())
    d = np.diff(t)
    print("rep %d: fill %.3f s  upload %.3f s  sweep_run (host, queueing) %.3f s  wait for the GPU %.3f s  download %.3f s  total %.3f s"
          % ((rep,) + tuple(d) + (t[-1] - t[0],)))
    info = {}
    t0 = time.perf_counter()
    sampling.sweep(like, P, device=torch.device("cuda:0"), info=info)
    print("       sampling.sweep wall %.3f s, device_ms %.1f" % (time.perf_counter() - t0, info["device_ms"]))
