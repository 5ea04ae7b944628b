"""theta-level parity on prior draws (VERDICT r01 "next" item 1).

Two layers:

* ``test_prior_*``: the committed fixture ``tests/golden/m2_prior.npz`` -- the reference's lnL, early-exit flag and
  signal marginals for 540 ST-U parameter vectors (the bench's first 512, 16 near the truth, 12 compact stars with
  two and three image orders) -- against the theta-level GPU path (embed + integrate + fold + likelihood).
* ``test_stage_attribution_live``: for the 64-row selection the reference build in ``oracle/_ref`` is run on the
  box (it is the checker, never the product), every integrator input is captured, and the GPU path is evaluated
  (a) on the reference's mesh and rays (the 1e-6 bar of BASELINE.json applies here: identical inputs),
  (b) with its own mesh and/or rays swapped in, so the theta-level difference is attributed to a stage,
  (c) for status equality with the reference's early exits.
"""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

KEV = 1.60217662e-16

# Measured on B200 (profiles/r02_theta_parity.txt); every bound below is <= 10x the measured maximum.
SHARED_MESH_LNL_ATOL = 1.0e-6          # BASELINE.json bar on identical integrator inputs
SAME_MESH_LNL_ATOL = 1.6e-5            # from theta, vectors whose member areas agree to 1e-11 (measured 1.6e-6 at
                                       # |lnL| = 2e6, 7.9e-13 relative: cell-to-cell redistribution below that)
THETA_LEVEL_REL = 3.0e-8               # |dlnL| / |lnL| where the reference's own mesh is off
THETA_LEVEL_ABS_NEAR_TRUTH = 3.0e-7    # |dlnL| for the 16 vectors around the truth (|lnL| ~ 4e4)
MESH_AREA_RTOL = 5.0e-9                # sum of GPU cell areas vs exact spot area
MARGINAL_RTOL = 1.0e-8                 # energy- and phase-summed pulse of each member


def _pipeline(max_batch):
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.pipeline import BatchedLikelihood
    m2 = np.load(os.path.join(GOLDEN, "m2_stu_nsx.npz"))
    matrix, edges = syn.nicer_like_response()[:2]
    pipe = BatchedLikelihood(member_component=[0, 1], max_rings=64, max_azi=64, n_rays=200,
                             energies=m2["t0_int0_energies"], leaves=m2["t0_int0_leaves"],
                             phases=m2["t0_int0_phases"], hot_atm_ext=2, hot_atmosphere=syn.nsx_like_table(),
                             image_order_limit=3, response=matrix, energy_edges=edges, counts=m2["counts"],
                             data_phases=np.linspace(0.0, 1.0, 33), exposure_time=syn.M2_EXPOSURE,
                             max_batch=max_batch)
    return pipe, m2["t0_int0_energies"]


def _report(tag, lnL, ref, mask):
    d = np.abs(lnL[mask] - ref[mask])
    rel = d / np.abs(ref[mask])
    k = int(np.argmax(d))
    print("%-46s n=%3d  max|d|=%.3e (lnL %.4e)  max|d|/|lnL|=%.3e  median|d|=%.3e"
          % (tag, mask.sum(), d.max(), ref[mask][k], rel.max(), np.median(d)))
    return d, rel


def _worst(lnL, ref, mask, fixture, rows, k=8):
    idx = np.flatnonzero(mask)
    d = np.abs(lnL[idx] - ref[idx])
    for j in np.argsort(-d)[:k]:
        i = idx[j]
        r = rows[i]
        print("    row %3d |d|=%.3e rel=%.2e lnL=%.4e rings=%s maxdefl=%.3f polar=%d R/r_g=%.3f"
              % (r, d[j], d[j] / abs(ref[i]), ref[i], fixture["n_rings"][r].tolist(),
                 fixture["max_deflection"][r].max(), fixture["polar"][r],
                 fixture["thetas"][r, 1] * 1e3 / (fixture["thetas"][r, 0] * 1476.62512)))


def test_prior_draws_status_and_lnL_from_theta():
    """540 parameter vectors, everything on the GPU from theta.

    (c) status: the reference early-exits (random value near llzero, default_background_marginalisation.pyx:
        677-684) on exactly the vectors where the GPU path reports status 11/12.
    (b) mesh: the sum of every member's GPU cell areas reproduces the exact spot area (40-digit quadrature in
        the fixture); wherever the GPU mesh and the reference build's mesh differ, the GPU mesh is the closer one.
        The reference's adaptive quadrature of the kinked boundary-cell integrands (mesh_tools.pyx:429-473) is
        good to its own epsrel = 1e-8 only, and misses slivers altogether.
    (a') lnL: on the vectors whose two meshes agree, |dlnL| stays below the bound; on the others the difference
        follows the reference's area error (reported, bounded relative to |lnL|)."""
    from xpsi_b200 import synthetic as syn
    d = np.load(os.path.join(GOLDEN, "m2_prior.npz"))
    thetas, ref, early = d["thetas"], d["lnL"], d["early_exit"]
    n = thetas.shape[0]
    pipe, energies = _pipeline(270)
    lnL = np.empty(n)
    status = np.empty(n, dtype=np.int32)
    sum_E = np.empty((n, 2, 100))
    sum_P = np.empty((n, 2, 128))
    area = np.empty((n, 2))
    for i in range(0, n, 270):
        blk = thetas[i:i + 270]
        lnL[i:i + len(blk)], status[i:i + len(blk)] = pipe.eval_spots(syn.m2_spot_batch(pipe, blk))
        flux = pipe.fetch(len(blk), folded=False, expected=False)[0] / (energies[None, :, None] * KEV)
        flux = flux.reshape(len(blk), 2, 128, 100)
        sum_E[i:i + len(blk)] = flux.sum(axis=2)
        sum_P[i:i + len(blk)] = flux.sum(axis=3)
        e = pipe.fetch_embed(len(blk))
        assert (e["n_rings"].reshape(-1, 2) == d["n_rings"][i:i + len(blk)]).all()
        area[i:i + len(blk)] = e["cellArea"].sum(axis=(1, 2)).reshape(-1, 2)
    codes, cnt = np.unique(status, return_counts=True)
    print("status counts", dict(zip(codes.tolist(), cnt.tolist())), " reference early exits", int(early.sum()))
    assert np.isin(status, (0, 11, 12)).all()
    assert ((status != 0) == early).all(), "early-exit sets differ: %r" % np.flatnonzero((status != 0) != early)
    # ---- mesh against the exact spot area ---------------------------------------------------------------
    exact = d["exact_area"]
    err_gpu = np.abs(area / exact - 1.0)
    err_ref = np.abs(d["area_sum"] / exact - 1.0)
    differ = np.abs(area / d["area_sum"] - 1.0) > 1.0e-11          # members whose two meshes are not the same
    print("sum of cell areas vs exact spot area: GPU max %.2e | reference max %.2e; meshes differ for %d of %d "
          "members, GPU closer to exact in %d of them"
          % (err_gpu.max(), err_ref.max(), differ.sum(), differ.size, (err_gpu[differ] < err_ref[differ]).sum()))
    # two members (a secondary spot ending 6e-4 rad short of the south pole) lose 6e-4 of their area to the
    # meshing scheme itself (the last ring's parallel, mesh.pyx:95-239) in both builds alike: not a quadrature matter
    scheme = err_ref < 1.0e-6
    print("members whose area the meshing scheme itself misses by > 1e-6: %d (GPU and reference agree there to %.1e)"
          % ((~scheme).sum(), np.abs(area / d["area_sum"] - 1.0)[~scheme].max() if (~scheme).any() else 0.0))
    assert err_gpu[scheme].max() < MESH_AREA_RTOL
    # below ~1e-9 both sums carry the scheme's own parallels (1000-node area table); above it the exact
    # area discriminates: every member the reference gets wrong by > 1e-8 the GPU mesh has 10x closer
    clear = differ & scheme & (err_ref > 1.0e-8)
    print("members with a reference area error > 1e-8: %d; GPU error there max %.2e" % (clear.sum(), err_gpu[clear].max()))
    assert clear.sum() >= 5 and (err_gpu[clear] < 0.1 * err_ref[clear]).all()
    assert (np.abs(area / d["area_sum"] - 1.0)[~scheme] < 1.0e-6).all()
    # ---- lnL ---------------------------------------------------------------------------------------------
    ok = ~early
    same = ok & ~differ.any(axis=1)
    nb = int(d["n_bench"])
    idx = np.arange(n)
    _report("GPU embed, all status-0 vectors", lnL, ref, ok)
    ds, rs = _report("  meshes agree (<1e-11 in area)", lnL, ref, same)
    dm, rm = _report("  meshes differ (reference area error)", lnL, ref, ok & ~same)
    dn, _ = _report("  near the truth", lnL, ref, ok & (idx >= nb) & (idx < nb + 16))
    _report("  compact stars (2-3 image orders)", lnL, ref, ok & (idx >= nb + 16))
    _report("  polar caps", lnL, ref, ok & d["polar"])
    _worst(lnL, ref, ok, d, idx)
    assert ds.max() < SAME_MESH_LNL_ATOL
    assert rm.max() < THETA_LEVEL_REL
    assert dn.max() < THETA_LEVEL_ABS_NEAR_TRUTH
    # pulse marginals of every member whose meshes agree, early exits included (they run the whole integrator)
    keep = ~differ
    def rel(got, want):                 # a member that is never visible has an identically zero signal in both
        scale = np.max(np.abs(want), axis=2, keepdims=True)
        dark = scale == 0.0
        assert (np.abs(got) * dark == 0.0).all(), "reference signal is dark, GPU signal is not"
        return (np.abs(got - want) / np.where(dark, 1.0, scale)).max(axis=2)
    eE, eP = rel(sum_E, d["flux_sum_E"]), rel(sum_P, d["flux_sum_P"])
    print("member pulse marginals (meshes agree): energy-summed rel err %.3e, phase-summed rel err %.3e; "
          "(meshes differ: %.3e, %.3e)" % (eE[keep].max(), eP[keep].max(), eE[differ].max(), eP[differ].max()))
    assert eE[keep].max() < MARGINAL_RTOL and eP[keep].max() < MARGINAL_RTOL


def _fill_from_reference(batch, b, r, mesh=None, rays=None, q0=None):
    """Member arrays of parameter vector ``b``: the reference's (default), or the GPU embed's ``mesh`` and/or
    ``rays`` (dict from ``fetch_embed``, rows q0, q0+1) swapped in."""
    batch.omega[b] = r["members"][0]["omega"]
    batch.inclination[b] = r["members"][0]["inclination"]
    batch.d_sq[b] = r["d_sq"]
    batch.phase_shifts[b] = r["phase_shifts"]
    for m, mem in enumerate(r["members"]):
        batch.set_member(b, m, mem["cellArea"], mem["theta"], mem["phi"], mem["radialCoords_of_parallels"],
                         mem["r_s_over_r"], mem["srcCellParams"], mem["deflection"], mem["cos_alpha"], mem["lag"],
                         mem["maxDeflection"], mem["cos_gammaArray"])
        q = b * batch.M + m
        R = mem["cellArea"].shape[0]
        if mesh is not None:
            e, qe = mesh, q0 + m
            assert e["n_rings"][qe] == R
            batch.cellArea[q] = e["cellArea"][qe]
            batch.phi[q] = e["phi"][qe]
            batch.theta[q] = e["theta"][qe]
            batch.radial[q] = e["radial"][qe]
            batch.r_s_over_r[q, :R] = mem["r_s"] / e["radial"][qe, :R]
            batch.srcParams[q] = e["srcParams"][qe]
            batch.cos_gamma[q] = e["cos_gamma"][qe]
        if rays is not None:
            e, qe = rays, q0 + m
            for k in ("deflection", "cos_alpha", "lag", "maxDeflection"):
                getattr(batch, k)[q] = e[k][qe]


def test_stage_attribution_live():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, GOLDEN)
    import ref_env
    if not ref_env.available():
        pytest.skip("oracle/_ref (reference build) is not on this box")
    import prior_ref
    from xpsi_b200 import synthetic as syn
    d = np.load(os.path.join(GOLDEN, "m2_prior.npz"))
    sel = d["sel64"]
    thetas = d["thetas"][sel]
    n = len(sel)
    assert n >= 64
    res = prior_ref.run_reference(thetas, full=True)
    ref = np.array([r["lnL"] for r in res])
    early = ref < -1.0e80
    # the live run reproduces the committed fixture (same build, same inputs)
    assert (early == d["early_exit"][sel]).all()
    assert np.max(np.abs(ref[~early] / d["lnL"][sel][~early] - 1.0)) < 1.0e-12
    pipe, energies = _pipeline(n)
    # theta level: everything embedded on the GPU
    lnL_gpu, st_gpu = pipe.eval_spots(syn.m2_spot_batch(pipe, thetas))
    emb = pipe.fetch_embed(n)
    for b, r in enumerate(res):
        for m, mem in enumerate(r["members"]):
            assert emb["n_rings"][2 * b + m] == mem["cellArea"].shape[0], (b, m)
    out = {}
    for tag, use_mesh, use_rays in (("reference mesh + reference rays", False, False),
                                    ("GPU mesh + reference rays", True, False),
                                    ("reference mesh + GPU rays", False, True)):
        batch = pipe.new_batch(n)
        for b, r in enumerate(res):
            _fill_from_reference(batch, b, r, emb if use_mesh else None, emb if use_rays else None, 2 * b)
        out[tag] = pipe(batch)
        if tag.startswith("reference mesh + reference"):
            flux = pipe.fetch(n, folded=False, expected=False)[0] / (energies[None, :, None] * KEV)
    ok = ~early
    print()
    for tag, (lnL, st) in out.items():
        assert ((st != 0) == early).all(), tag
        assert np.isin(st, (0, 11, 12)).all(), tag
        _report(tag, lnL, ref, ok)
    assert ((st_gpu != 0) == early).all()
    _report("GPU mesh + GPU rays (theta level)", lnL_gpu, ref, ok)
    for tag in ("GPU mesh + reference rays", "reference mesh + GPU rays"):
        print("  worst rows,", tag)
        _worst(out[tag][0], ref, ok, d, sel)
    # (a) identical integrator inputs: the 1e-6 absolute bar
    lnL_a = out["reference mesh + reference rays"][0]
    da = np.abs(lnL_a[ok] - ref[ok])
    assert da.max() < SHARED_MESH_LNL_ATOL, "shared-mesh lnL differs by %.3e" % da.max()
    # pulse signals of every member on identical inputs: 1e-8 relative
    worst = 0.0
    for b, r in enumerate(res):
        for m, mem in enumerate(r["members"]):
            scale = np.max(np.abs(mem["flux"]), axis=1, keepdims=True)
            scale[scale == 0.0] = 1.0
            worst = max(worst, float(np.max(np.abs(flux[2 * b + m] - mem["flux"]) / scale)))
    print("pulse signals on identical inputs: max rel err %.3e over %d member signals" % (worst, 2 * n))
    assert worst < 1.0e-8
    # (b) mesh: where the two meshes differ the GPU one is closer to the exact spot area
    for b, r in enumerate(res):
        for m, mem in enumerate(r["members"]):
            ex = d["exact_area"][sel[b], m]
            a_gpu, a_ref = emb["cellArea"][2 * b + m].sum(), mem["cellArea"].sum()
            if abs(a_gpu / a_ref - 1.0) > 1.0e-11:
                print("    row %3d member %d: mesh area vs exact, GPU %.2e  reference %.2e"
                      % (sel[b], m, a_gpu / ex - 1.0, a_ref / ex - 1.0))
                assert abs(a_gpu / ex - 1.0) <= abs(a_ref / ex - 1.0) or abs(a_ref / ex - 1.0) > 1.0e-6
    # (b) attribution: the theta-level difference is carried by the embed stages, not by the integrator
    d_mesh = np.abs(out["GPU mesh + reference rays"][0][ok] - ref[ok])
    d_rays = np.abs(out["reference mesh + GPU rays"][0][ok] - ref[ok])
    d_full = np.abs(lnL_gpu[ok] - ref[ok])
    print("attribution (max over %d vectors): integrator %.2e | mesh swap %.2e | rays swap %.2e | both %.2e"
          % (ok.sum(), da.max(), d_mesh.max(), d_rays.max(), d_full.max()))
    assert (d_full / np.abs(ref[ok])).max() < THETA_LEVEL_REL
    assert d_rays.max() < 1.0e-6
