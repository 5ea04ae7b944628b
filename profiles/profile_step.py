#!/usr/bin/env python
"""One profiled step of the bench workload (batch --batch distinct parameter vectors, embed + four stages) for ncu.

    ncu --profile-from-start off ... python profiles/profile_step.py --batch 512

Warm-up steps run outside the cudaProfilerStart/Stop window; exactly one step is inside it.
Driven by profiles/make_profiles.sh, which turns the captures into profiles/r02_kernels.json (read by bench.py).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--warm", type=int, default=2)
a = ap.parse_args()

import torch  # noqa: E402  (cudaProfilerStart/Stop only)
import bench  # noqa: E402
from xpsi_b200 import synthetic as syn  # noqa: E402

w = bench.load_workload()
pipe = bench.make_pipeline(w, a.batch)
P = syn.m2_bench_thetas(0, a.batch * (a.warm + 1))
pipe.sweep_upload(syn.m2_spot_batch(pipe, P))
pipe.sweep_run(0, a.batch * a.warm)
torch.cuda.synchronize()
torch.cuda.profiler.start()
pipe.sweep_run(a.batch * a.warm, a.batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
lnL, st = pipe.sweep_download(a.batch * a.warm, a.batch)
print("profiled step: batch", a.batch, "status-0 rows", int((st == 0).sum()))
