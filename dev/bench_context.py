#!/usr/bin/env python
"""Which part of bench.py's sequence changes the time of the 1e5 sweep that follows it?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from xpsi_b200 import _lib, sampling, synthetic as syn  # noqa: E402
from xpsi_b200.likelihood import Likelihood  # noqa: E402

B, K, W = 512, 10, 3
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
_lib.check(_lib.lib.xpsi_b200_set_device(0))
w = bench.load_workload()
pipe = bench.make_pipeline(w, B)
like = Likelihood(pipe, lambda p, P: syn.m2_spot_batch(p, P), prior=None)
P_all = syn.m2_bench_thetas(0, 100000)
stream = torch.cuda.ExternalStream(_lib.lib.xpsi_b200_stream(), device=dev)


def sweep(tag):
    sampling.sweep(like, P_all[:2 * B], device=dev)
    torch.cuda.synchronize()
    info = {}
    t0 = time.perf_counter()
    sampling.sweep(like, P_all, device=dev, info=info)
    print("%-40s sweep wall %.3f s (device %.1f ms)" % (tag, time.perf_counter() - t0, info["device_ms"]), flush=True)


sweep("fresh pipeline")
peak = np.zeros(1)
_lib.check(_lib.lib.xpsi_b200_fp64_peak_tflops(_lib.dptr(peak)))
sweep("after the DFMA peak kernel")
thetas_all = syn.m2_bench_thetas(0, (W + K) * B)
pipe.sweep_upload(syn.m2_spot_batch(pipe, thetas_all))
pipe.sweep_run(0, W * B)
with torch.cuda.stream(stream):
    pipe.sweep_run(W * B, K * B)
    d_lnL, d_st = pipe.sweep_device_results()
    t = torch.cat([torch.as_tensor(d_lnL, device=dev)[W * B:], torch.as_tensor(d_st, device=dev)[W * B:].to(torch.float64)])
torch.cuda.synchronize()
sweep("after the timed steps")
s = bench.ClockSampler(0)
s.start()
time.sleep(1.0)
print(s.stop())
sweep("after a ClockSampler")
pipe.sweep_upload(syn.m2_spot_batch(pipe, thetas_all))
pipe.count_work(True)
for k in range(K):
    pipe.sweep_run((W + k) * B, B)
    pipe.stage_ms()
print(pipe.count_work(False))
sweep("after the counting pass")
