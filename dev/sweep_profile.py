#!/usr/bin/env python
"""Per-block stage times along a long sweep (does the cost per block drift with the block index?).
    python dev/sweep_profile.py [--batch 512] [--blocks 194] [--sync-every 1]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--blocks", type=int, default=194)
a = ap.parse_args()
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from xpsi_b200 import _lib, synthetic as syn  # noqa: E402

w = bench.load_workload()
pipe = bench.make_pipeline(w, a.batch)
P = syn.m2_bench_thetas(0, a.batch * (a.blocks + 2))
pipe.sweep_upload(syn.m2_spot_batch(pipe, P))
pipe.sweep_run(0, a.batch * 2)
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(_lib.lib.xpsi_b200_stream())
rows = []
for b in range(a.blocks):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        pipe.sweep_run(a.batch * (2 + b), a.batch)
        e1.record()
    e1.synchronize()
    s = pipe.stage_ms()
    rows.append([e0.elapsed_time(e1)] + [s[k] for k in ("embed", "integrate", "energy", "fold", "marginal", "flux_kernel")])
rows = np.array(rows)
print("blocks      total   embed  integr  energy    fold  margin    flux")
for i in range(0, a.blocks, 16):
    print("%3d-%3d  " % (i, min(i + 16, a.blocks) - 1) + " ".join("%7.3f" % v for v in rows[i:i + 16].mean(axis=0)))
print("all      " + " ".join("%7.3f" % v for v in rows.mean(axis=0)))
print("max block", int(rows[:, 0].argmax()), rows[rows[:, 0].argmax()])
# the same rows as ONE call (what sampling.sweep does)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record()
    pipe.sweep_run(a.batch * 2, a.batch * a.blocks)
    e1.record()
e1.synchronize()
print("one sweep_run over the same %d blocks: %.3f ms per block" % (a.blocks, e0.elapsed_time(e1) / a.blocks))
