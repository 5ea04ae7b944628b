#!/usr/bin/env python
"""Stage times of the bench workload for the library named by XPSI_B200_LIB (default: the in-tree build).
    python dev/time_step.py [--batch 512] [--blocks 6]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--blocks", type=int, default=6)
ap.add_argument("--deterministic", action="store_true")
ap.add_argument("--dump", default=None, help="save lnL / status of the evaluated rows (npz)")
a = ap.parse_args()
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from xpsi_b200 import _lib, synthetic as syn  # noqa: E402

w = bench.load_workload()
pipe = bench.make_pipeline(w, a.batch)
if a.deterministic:
    pipe.set_deterministic(True)
P = syn.m2_bench_thetas(0, a.batch * (a.blocks + 2))
pipe.sweep_upload(syn.m2_spot_batch(pipe, P))
pipe.sweep_run(0, a.batch * 2)
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(_lib.lib.xpsi_b200_stream())
acc = {}
tot = 0.0
for b in range(a.blocks):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        pipe.sweep_run(a.batch * (2 + b), a.batch)
        e1.record()
    e1.synchronize()
    tot += e0.elapsed_time(e1)
    for k, v in pipe.stage_ms().items():
        acc[k] = acc.get(k, 0.0) + v
lnL, st = pipe.sweep_download(0, a.batch * (a.blocks + 2))
if a.dump:
    np.savez(a.dump, lnL=lnL, status=st)
print(os.environ.get("XPSI_B200_LIB", "in-tree"), "batch", a.batch, "ms/block %.3f" % (tot / a.blocks),
      "evals/s %.0f" % (a.batch * a.blocks / tot * 1e3), {k: round(v / a.blocks, 3) for k, v in acc.items()},
      "lnL checksum %.10e" % float(np.nansum(np.where(st == 0, lnL, 0.0))))
