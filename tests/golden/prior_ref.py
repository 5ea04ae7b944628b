"""Run the reference build (oracle/_ref) on lists of ST-U parameter vectors (TEST INFRASTRUCTURE).

Used by ``make_golden_prior.py`` (committed fixture ``m2_prior.npz``), by the live stage-attribution parity test
(``tests/test_theta_parity.py``) and by nothing under ``xpsi_b200/``.  Each worker process builds the M2 model
of ``make_golden.build_m2`` once, evaluates ``likelihood(theta, force=True)`` per row and returns, per theta,
the reference's lnL, its early-exit flag, per-member mesh summaries and either compact signal marginals or
(``full=True``) every array that crosses the integrator boundary.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

FULL_KEYS = ("cellArea", "theta", "phi", "radialCoords_of_parallels", "r_s_over_r", "srcCellParams",
             "deflection", "cos_alpha", "lag", "maxDeflection", "cos_gammaArray")


def _worker(args):
    thetas, full = args
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    out = []
    with contextlib.redirect_stdout(io.StringIO()):
        import ref_env
        ref_env.import_reference()
        import make_golden as mg
        rec = mg.Recorder()
        rec.wrap_signal_module()
        m2 = np.load(os.path.join(HERE, "m2_stu_nsx.npz"))
        like, signal, instrument, hots = mg.build_m2(rec, m2["counts"])
        names = mg.INTEGRATE_ARGS
        for th in thetas:
            rec.clear()
            lnL = float(like(list(th), force=True))
            r = dict(theta=np.array(th, dtype=np.float64), lnL=lnL, early_exit=bool(lnL < -1.0e80),
                     d_sq=float(like.star.spacetime.d_sq), members=[])
            for call in rec.calls["integrate"]:
                a = dict(zip(names, call[0]))
                flux = np.asarray(call[2][1])
                mem = dict(n_rings=int(a["cellArea"].shape[0]),
                           max_deflection=float(np.max(a["maxDeflection"])),
                           area_sum=float(np.sum(a["cellArea"])),
                           polar=bool(np.any(np.asarray(a["theta"])[:, 0] <= 0.0) or a["cellArea"].shape[0] !=
                                      a["cellArea"].shape[1]),
                           flux_sum_E=flux.sum(axis=0), flux_sum_P=flux.sum(axis=1))
                if full:
                    for k in FULL_KEYS:
                        mem[k] = np.array(a[k])
                    mem["flux"] = flux.copy()
                    for k in ("R", "omega", "r_s", "inclination"):
                        mem[k] = float(a[k])
                r["members"].append(mem)
            if rec.calls["marginal"]:
                ma, mres = rec.calls["marginal"][0]
                r["phase_shifts"] = np.array(ma[5])
                if full:
                    r["components"] = [np.array(c) for c in ma[3]]
            out.append(r)
    return out


def exact_spot_area(args):
    """Area of a circular spot (colatitude ``colat``, angular radius ``rho``) on the oblate AlGendy-Morsink
    surface (mesh_tools.pyx:18-98) by 40-digit quadrature; independent of any mesh."""
    epsilon, zeta, R_eq, colat, rho = args
    import mpmath as mp
    mp.mp.dps = 40
    eps, zeta, Req = mp.mpf(epsilon), mp.mpf(zeta), mp.mpf(R_eq)
    tc, rho = mp.mpf(colat), mp.mpf(rho)
    k = mp.mpf("-0.788") + mp.mpf("1.030") * zeta

    def g(t):
        mu = mp.cos(t)
        r = 1 + eps * k * mu * mu
        f = (-2 * eps * k * mu * mp.sqrt(1 - mu * mu)) / (r * mp.sqrt(1 - 2 * zeta / r))
        c = (mp.cos(rho) - mp.cos(tc) * mu) / (mp.sin(tc) * mp.sin(t))
        phi_b = mp.pi if c <= -1 else (mp.mpf(0) if c >= 1 else mp.acos(c))
        return 2 * phi_b * r * r * mp.sqrt(1 + f * f) * mp.sin(t)
    lo, hi = max(tc - rho, mp.mpf(0)), min(tc + rho, mp.pi)
    pts = [lo]
    for brk in (rho - tc, 2 * mp.pi - tc - rho):            # where a cap starts to cover whole parallels
        if lo < brk < hi:
            pts.append(brk)
    pts.append(hi)
    full = []
    for a, b in zip(pts[:-1], pts[1:]):
        full += [a + (b - a) * mp.mpf(j) / 8 for j in range(8)]
    full.append(hi)
    return float(Req * Req * mp.quad(g, full))


def run_reference(thetas, full=False, procs=None):
    """Evaluate the reference on ``thetas[n, 11]`` in ``procs`` processes (threads=1 each); returns a list of
    per-theta dicts in input order."""
    import multiprocessing as mp
    thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    procs = min(procs or len(os.sched_getaffinity(0)), len(thetas))
    chunks = [thetas[i::procs] for i in range(procs)]
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(_worker, [(c, full) for c in chunks])
    out = [None] * len(thetas)
    for i, rows in enumerate(res):
        for j, r in enumerate(rows):
            out[i + j * procs] = r
    return out
