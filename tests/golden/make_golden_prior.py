#!/usr/bin/env python
"""Golden fixture ``m2_prior.npz``: the reference's lnL / early-exit flag / mesh summaries / signal marginals
for (i) the first 512 rows of the bench's ST-U parameter-vector list (``synthetic.m2_bench_thetas``) and
(ii) 16 parameter vectors scattered around the truth and (iii) 12 stars close to R = 3 r_g (two and three image
orders), plus a 64-row selection ``sel64`` for the stage-attribution parity test (12+ polar caps, the 12 compact
stars, the 16 near the truth, the rest the head of the bench list).

Run here (needs oracle/_ref):   python tests/golden/make_golden_prior.py
"""
import math
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from xpsi_b200 import synthetic as syn  # noqa: E402
import prior_ref  # noqa: E402

N_BENCH, N_NEAR, N_COMPACT = 512, 16, 12


def main():
    thetas = np.vstack([syn.m2_bench_thetas(0, N_BENCH), syn.m2_near_truth_thetas(N_NEAR), syn.m2_compact_thetas()])
    t0 = time.time()
    res = prior_ref.run_reference(thetas)
    print("reference: %d evaluations in %.0f s" % (len(res), time.time() - t0))
    n = len(res)
    lnL = np.array([r["lnL"] for r in res])
    early = np.array([r["early_exit"] for r in res])
    n_rings = np.array([[m["n_rings"] for m in r["members"]] for r in res], dtype=np.int32)
    max_defl = np.array([[m["max_deflection"] for m in r["members"]] for r in res])
    sum_E = np.array([[m["flux_sum_E"] for m in r["members"]] for r in res])
    sum_P = np.array([[m["flux_sum_P"] for m in r["members"]] for r in res])
    area_sum = np.array([[m["area_sum"] for m in r["members"]] for r in res])
    # exact spot areas (40-digit quadrature, no mesh involved): the yardstick for mesh differences
    import multiprocessing as mp
    jobs = []
    for th in thetas:
        st = syn.SpacetimeScalars(th[0], th[1], th[2], th[3], syn.M2_FREQUENCY)
        jobs += [(st.epsilon, st.zeta, st.R, th[5], th[6]), (st.epsilon, st.zeta, st.R, th[9], th[10])]
    with mp.get_context("spawn").Pool(len(os.sched_getaffinity(0))) as pool:
        exact_area = np.array(pool.map(prior_ref.exact_spot_area, jobs, chunksize=8)).reshape(n, 2)
    rel = area_sum / exact_area - 1.0
    print("reference mesh: sum of cell areas vs exact spot area, |rel| max %.2e, members above 1e-9: %d of %d"
          % (np.abs(rel).max(), (np.abs(rel) > 1e-9).sum(), rel.size))
    polar = (thetas[:, 5] - thetas[:, 6] < 0.0) | (thetas[:, 9] + thetas[:, 10] > math.pi)
    n_img = np.minimum(3, np.ceil(max_defl.max(axis=1) / math.pi)).astype(np.int32)
    idx = np.arange(N_BENCH)
    sel = list(range(24))
    sel += [i for i in idx[polar[:N_BENCH]] if i not in sel][:12]
    sel += list(range(N_BENCH, N_BENCH + N_NEAR + N_COMPACT))
    # the four bench rows whose reference mesh is furthest from the exact spot area
    worst = np.argsort(-np.abs(rel[:N_BENCH]).max(axis=1))
    sel += [int(i) for i in worst if not early[i]][:4]
    sel = np.array(sorted(set(sel)), dtype=np.int32)
    print("selection: %d rows, %d polar, %d dual-image, %d triple-image, %d near truth, %d early exits"
          % (len(sel), polar[sel].sum(), (n_img[sel] == 2).sum(), (n_img[sel] >= 3).sum(),
             ((sel >= N_BENCH) & (sel < N_BENCH + N_NEAR)).sum(), early[sel].sum()))
    print("whole set: early exits %d / %d, |lnL| range of the rest %.3e .. %.3e"
          % (early.sum(), n, np.abs(lnL[~early]).min(), np.abs(lnL[~early]).max()))
    np.savez_compressed(os.path.join(HERE, "m2_prior.npz"), thetas=thetas, lnL=lnL, early_exit=early,
                        n_rings=n_rings, max_deflection=max_defl, area_sum=area_sum, exact_area=exact_area, polar=polar, n_img=n_img,
                        flux_sum_E=sum_E, flux_sum_P=sum_P, sel64=sel, n_bench=np.asarray(N_BENCH))
    print("m2_prior.npz", os.path.getsize(os.path.join(HERE, "m2_prior.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
