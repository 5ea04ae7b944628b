"""How many leaf intervals does a ring's cell walk touch per output phase?  (sizing the flux kernel's stage 3)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from xpsi_b200 import synthetic as syn
w = bench.load_workload()
B = 256
pipe = bench.make_pipeline(w, B)
P = syn.m2_bench_thetas(0, B)
pipe.embed_spots(syn.m2_spot_batch(pipe, P))
e = pipe.fetch_embed(B)
area, phi, nr = e["cellArea"], e["phi"], e["n_rings"]
cells, ivals = [], []
for q in range(2 * B):
    for r in range(nr[q]):
        rad = area[q, r] > 0
        n = int(rad.sum())
        if n == 0:
            continue
        ph = phi[q, r][rad]
        span = ph.max() - ph.min()
        cells.append(n)
        ivals.append(span / (2 * np.pi / 99) + 1)
cells, ivals = np.array(cells), np.array(ivals)
print("lit rings", len(cells), "rings per member", nr.mean())
print("cells per ring: mean %.1f p50 %.0f p90 %.0f max %d" % (cells.mean(), np.median(cells), np.percentile(cells, 90), cells.max()))
print("intervals per (ring, phase): mean %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f; > 24: %.2f%%"
      % (ivals.mean(), np.median(ivals), np.percentile(ivals, 90), np.percentile(ivals, 99), ivals.max(), 100 * (ivals > 24).mean()))
print("sum intervals / sum cells = %.3f" % (ivals.sum() / cells.sum()))
h, edges = np.histogram(ivals, bins=[0, 2, 4, 6, 8, 12, 16, 24, 32, 48, 100])
print("histogram", dict(zip(edges[1:].tolist(), h.tolist())))
