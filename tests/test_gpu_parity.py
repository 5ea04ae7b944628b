"""GPU parity tests: the CUDA hot path (called through the C ABI via the
reference-shaped Python callables) against golden vectors recorded from the
reference build (tests/golden/make_golden.py).

Tolerances are BASELINE.json's: pulse signals 1e-8 relative, lnL 1e-6 absolute.
"""
import os

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

PULSE_RTOL = 1.0e-8
LNL_ATOL = 1.0e-6
THETA_GOLDEN_LNL_ATOL = 5.0e-8     # lnL from theta (GPU embed) on the golden parameter vectors: 10x the measured 4.7e-9


def _integrate_args(d, prefix, atmosphere):
    g = lambda k: d[prefix + k]
    iol = int(g("image_order_limit")) if (prefix + "image_order_limit") in d.files else None
    return (1, float(g("R")), float(g("omega")), float(g("r_s")), float(g("inclination")),
            g("cellArea"), g("radialCoords_of_parallels"), g("r_s_over_r"), g("theta"), g("phi"),
            g("srcCellParams"), g("CELL_RADIATES"), None, int(g("numRays")), g("deflection"),
            g("cos_alpha"), g("lag"), g("maxDeflection"), g("cos_gammaArray"), g("energies"),
            g("leaves"), g("phases"), atmosphere, (), int(g("hot_atm_ext")), 1, int(g("beam_opt")), iol)


def _pulse_err(out, ref):
    """elementwise error relative to the largest value of the same energy row"""
    scale = np.max(np.abs(ref), axis=1, keepdims=True)
    scale[scale == 0.0] = 1.0
    return float(np.max(np.abs(out - ref) / scale))


def test_c1_integrate_blackbody(c1):
    from xpsi_b200.cellmesh.integrator_for_azimuthal_invariance import integrate
    status, flux = integrate(*_integrate_args(c1, "int0_", ()))
    assert status == 0
    assert flux.shape == c1["int0_flux"].shape
    err = _pulse_err(flux, c1["int0_flux"])
    print("C1 flux max rel err", err, "global", rel_err(flux, c1["int0_flux"]))
    assert err < PULSE_RTOL


def test_c1_energy_integrator(c1):
    from xpsi_b200.tools import energy_integrator
    signal = c1["int0_flux"] / c1["d_sq"]
    out = energy_integrator(1, signal, c1["eint_log10_energies"], c1["eint_log10_edges"])
    ref = c1["eint0_out"]
    assert out.shape == ref.shape
    err = rel_err(out, ref)
    print("C1 energy_integrator rel err", err)
    assert err < PULSE_RTOL


def test_c1_fold(c1):
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.instrument import Instrument
    matrix = syn.c1_response()[0]
    inst = Instrument(matrix)
    out = inst(c1["eint0_out"], (0, matrix.shape[1]), (0, matrix.shape[0]))
    ref = c1["marg_components_0"]
    err = rel_err(out, ref)
    print("C1 fold rel err", err)
    assert err < 1.0e-13


def _marginal(d, prefix):
    from xpsi_b200.likelihoods import eval_marginal_likelihood
    n = int(d[prefix + "n_components"])
    comps = tuple(d["%scomponents_%d" % (prefix, i)] for i in range(n))
    cph = tuple(d["%scomponent_phases_%d" % (prefix, i)] for i in range(n))
    return eval_marginal_likelihood(float(d[prefix + "exposure_time"]), d[prefix + "phases"],
                                    d[prefix + "counts"], comps, cph, d[prefix + "phase_shifts"],
                                    d[prefix + "precomp"], d[prefix + "support"],
                                    int(d[prefix + "workspace_intervals"]), float(d[prefix + "epsabs"]),
                                    float(d[prefix + "epsrel"]), float(d[prefix + "epsilon"]),
                                    float(d[prefix + "sigmas"]), float(d[prefix + "llzero"]))


def test_c1_marginal_likelihood(c1):
    lnL, star, mcl, mcl_s = _marginal(c1, "marg_")
    print("C1 lnL", lnL, "ref", float(c1["marg_lnL"]), "diff", lnL - float(c1["marg_lnL"]))
    assert abs(lnL - float(c1["marg_lnL"])) < LNL_ATOL
    assert rel_err(star, c1["marg_expected_counts"]) < PULSE_RTOL
    assert rel_err(mcl, c1["marg_mcl_background"]) < 1.0e-6
    # the reference's own published known answer (xpsi/tests/test_likelihood.py:134, rtol 1e-5)
    assert abs(lnL - float(c1["known_answer_lnL"])) < 1.0e-5 * abs(float(c1["known_answer_lnL"]))


def test_c1_precomputation(c1):
    from xpsi_b200.likelihoods import precomputation
    out = precomputation(c1["marg_counts"].astype(np.int32))
    assert np.max(np.abs(out - c1["marg_precomp"])) < 1.0e-9


def test_c1_chain_end_to_end(c1):
    """integrate -> /d_sq -> energy_integrator -> fold -> marginal, all on the GPU,
    against the reference's lnL."""
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.cellmesh.integrator_for_azimuthal_invariance import integrate
    from xpsi_b200.instrument import fold
    from xpsi_b200.likelihoods import eval_marginal_likelihood
    from xpsi_b200.tools import energy_integrator
    status, flux = integrate(*_integrate_args(c1, "int0_", ()))
    assert status == 0
    integrated = energy_integrator(1, flux / c1["d_sq"], c1["eint_log10_energies"], c1["eint_log10_edges"])
    matrix = syn.c1_response()[0]
    folded = fold(matrix, integrated, (0, matrix.shape[1]), (0, matrix.shape[0]))
    d, p = c1, "marg_"
    lnL = eval_marginal_likelihood(float(d[p + "exposure_time"]), d[p + "phases"], d[p + "counts"],
                                   (folded,), (d[p + "component_phases_0"],), d[p + "phase_shifts"],
                                   d[p + "precomp"], d[p + "support"], 1000, 0.0, 1e-8, 1e-3, 10.0, -1e90)[0]
    print("C1 chain lnL", lnL, "diff", lnL - float(c1["lnL_total"]))
    assert abs(lnL - float(c1["lnL_total"])) < LNL_ATOL


@pytest.mark.parametrize("t,m", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_m2_integrate_num4d(m2, t, m):
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.cellmesh.integrator_for_azimuthal_invariance import integrate
    table = test_m2_integrate_num4d.__dict__.setdefault("table", syn.nsx_like_table())
    prefix = "t%d_int%d_" % (t, m)
    status, flux = integrate(*_integrate_args(m2, prefix, table))
    assert status == 0
    ref = m2[prefix + "flux"]
    err = _pulse_err(flux, ref)
    print("M2 theta %d member %d flux rel err" % (t, m), err)
    assert err < PULSE_RTOL


def test_m2_marginal_likelihood(m2):
    for t in range(int(m2["n_theta"])):
        p = "t%d_marg_" % t
        lnL = _marginal(m2, p)[0]
        print("M2 theta", t, "lnL", lnL, "diff", lnL - float(m2[p + "lnL"]))
        assert abs(lnL - float(m2[p + "lnL"])) < LNL_ATOL


def test_m2_energy_integrator_and_fold(m2):
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.instrument import fold
    from xpsi_b200.tools import energy_integrator
    signal = m2["t0_int0_flux"] / m2["t0_d_sq"]
    out = energy_integrator(1, signal, m2["t0_eint_log10_energies"], m2["t0_eint_log10_edges"])
    assert rel_err(out, m2["t0_eint0_out"]) < PULSE_RTOL
    matrix = syn.nicer_like_response()[0]
    folded = fold(matrix, m2["t0_eint0_out"], (0, matrix.shape[1]), (0, matrix.shape[0]))
    err = rel_err(folded, m2["t0_marg_components_0"])
    print("M2 fold rel err", err)
    assert err < 1.0e-12


def _m2_pipeline(m2, max_batch=8):
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.pipeline import BatchedLikelihood
    matrix, edges, channels, ch_edges = syn.nicer_like_response()
    return BatchedLikelihood(member_component=[0, 1], max_rings=64, max_azi=64, n_rays=200,
                             energies=m2["t0_int0_energies"], leaves=m2["t0_int0_leaves"],
                             phases=m2["t0_int0_phases"], hot_atm_ext=2,
                             hot_atmosphere=syn.nsx_like_table(), image_order_limit=3,
                             response=matrix, energy_edges=edges, counts=m2["counts"],
                             data_phases=np.linspace(0.0, 1.0, 33), exposure_time=syn.M2_EXPOSURE,
                             max_batch=max_batch)


def fill_m2_batch(m2, batch, order):
    for b, t in enumerate(order):
        batch.omega[b] = m2["t%d_int0_omega" % t]
        batch.inclination[b] = m2["t%d_int0_inclination" % t]
        batch.d_sq[b] = m2["t%d_d_sq" % t]
        batch.phase_shifts[b] = m2["t%d_marg_phase_shifts" % t]
        for m in range(2):
            g = lambda k: m2["t%d_int%d_%s" % (t, m, k)]
            batch.set_member(b, m, g("cellArea"), g("theta"), g("phi"), g("radialCoords_of_parallels"),
                             g("r_s_over_r"), g("srcCellParams"), g("deflection"), g("cos_alpha"),
                             g("lag"), g("maxDeflection"), g("cos_gammaArray"))


def test_m2_batched_pipeline(m2):
    pipe = _m2_pipeline(m2)
    order = [0, 1, 1, 0, 0]
    batch = pipe.new_batch(len(order))
    fill_m2_batch(m2, batch, order)
    lnL, status = pipe(batch)
    print("pipeline lnL", lnL, "status", status, "stages", pipe.stage_ms())
    assert (status == 0).all()
    for b, t in enumerate(order):
        ref = float(m2["t%d_lnL_total" % t])
        print("  theta", t, "lnL", lnL[b], "ref", ref, "diff", lnL[b] - ref)
        assert abs(lnL[b] - ref) < LNL_ATOL
    flux, folded, expected = pipe.fetch(len(order))
    for b, t in enumerate(order):
        for m in range(2):
            ref = m2["t%d_int%d_flux" % (t, m)]
            got = flux[b * 2 + m] / (m2["t0_int0_energies"][:, None] * 1.60217662e-16)
            assert _pulse_err(got, ref) < PULSE_RTOL
            assert rel_err(folded[b, m], m2["t%d_marg_components_%d" % (t, m)]) < PULSE_RTOL
        assert rel_err(expected[b], m2["t%d_marg_expected_counts" % t]) < PULSE_RTOL
    # identical inputs in different batch slots give identical answers up to the
    # order of the fp64 ring reduction
    assert abs(lnL[0] - lnL[3]) < 1e-7 and abs(lnL[1] - lnL[2]) < 1e-7


def test_signal_tools_against_reference_golden():
    """tools.phase_integrator / phase_interpolator / energy_interpolator (row a12)."""
    import os
    from conftest import GOLDEN
    from xpsi_b200.tools import energy_interpolator, phase_integrator, phase_interpolator
    d = np.load(os.path.join(GOLDEN, "tools.npz"))
    tag = lambda s: ("%+.2f" % s).replace(".", "p").replace("+", "P").replace("-", "M")
    for shift in d["shifts"]:
        out = phase_integrator(1000.0, d["edges"], d["pulse"], d["sig_phases"], float(shift))
        assert out.shape == d["pint_" + tag(shift)].shape
        assert rel_err(out, d["pint_" + tag(shift)]) < PULSE_RTOL
        out = phase_interpolator(d["new_phases"], d["sig_phases"], d["pulse"], float(shift))
        assert rel_err(out, d["pitp_" + tag(shift)]) < PULSE_RTOL
    out = energy_interpolator(1, d["flux"], d["log10E"], d["new_E"])
    assert out.shape == d["eitp"].shape
    print("energy_interpolator rel err", rel_err(out, d["eitp"]))
    assert rel_err(out, d["eitp"]) < PULSE_RTOL
    assert rel_err(energy_interpolator(1, d["flux_neg"], d["log10E"], d["new_E"]), d["eitp_neg"]) < PULSE_RTOL


def test_signal_tools_against_oracle_on_seeded_inputs():
    """Same three tools vs the CPU oracle on seeded ragged inputs (odd sizes, both interpolants)."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle as orc
    from xpsi_b200 import tools
    rng = np.random.default_rng(7)
    try:
        for kind in ("Akima", "Steffen"):
            tools.set_phase_interpolant(kind)
            for n_nodes, n_rows, n_bins in ((5, 1, 1), (33, 7, 13), (101, 3, 32)):
                ph = np.linspace(0.0, 1.0, n_nodes)
                sig = np.abs(np.sin(2 * np.pi * ph)[None, :] * rng.uniform(0.5, 2.0, (n_rows, 1)) + rng.uniform(0, .3, (n_rows, n_nodes)))
                sig[:, -1] = sig[:, 0]
                edges = np.linspace(0.0, 1.0, n_bins + 1)
                shift = float(rng.uniform(-0.5, 0.5))
                a = tools.phase_integrator(10.0, edges, sig, ph, shift)
                b = orc.phase_integrator(10.0, edges, sig, ph, shift, phase_interpolant=kind)
                assert rel_err(a, b) < 1e-12
                newp = np.sort(rng.uniform(0, 1, 17))
                a = tools.phase_interpolator(newp, ph, sig, shift, allow_negative=1)
                b = orc.phase_interpolator(newp, ph, sig, shift, allow_negative=1, phase_interpolant=kind)
                assert rel_err(a, b) < 1e-12
    finally:
        tools.set_phase_interpolant("Akima")


def _tinv_args(d, prefix, atmosphere):
    g = lambda k: d[prefix + k]
    return (1, float(g("R")), float(g("omega")), float(g("r_s")), float(g("inclination")), int(g("sqrt_numPix")),
            float(g("cellArea")), g("radialCoords_of_parallels"), g("r_s_over_r"), g("theta"), g("phi"),
            g("srcCellParams"), int(g("numRays")), g("deflection"), g("cos_alpha"), g("maxDeflection"),
            g("cos_gammaArray"), g("energies"), atmosphere, int(g("atm_ext")), int(g("image_order_limit")))


def test_time_invariant_integrator_everywhere_and_elsewhere():
    """integrator_for_time_invariance (row a3): Everywhere BB / Num4D and Elsewhere Num4D."""
    import os
    from conftest import GOLDEN
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.cellmesh.integrator_for_time_invariance import integrate
    table = syn.nsx_like_table()
    ev = np.load(os.path.join(GOLDEN, "m4_everywhere.npz"))
    el = np.load(os.path.join(GOLDEN, "m4_elsewhere.npz"))
    for d, prefix, atm in ((ev, "bb_", ()), (ev, "num4d_", table), (el, "else_", table)):
        status, flux = integrate(*_tinv_args(d, prefix, atm))
        assert status == 0
        err = float(np.max(np.abs(flux - d[prefix + "flux"]) / np.max(np.abs(d[prefix + "flux"]))))
        el_err = float(np.max(np.abs(flux / d[prefix + "flux"] - 1.0)))
        print("time-invariant", prefix, "rel err", err, "elementwise", el_err)
        assert err < PULSE_RTOL
    # a ring whose cells do NOT share (T, g) takes the per-cell 4-D path
    d = ev
    par = d["num4d_srcCellParams"].copy()
    par[:, ::2, 0] -= 0.05
    args = list(_tinv_args(d, "num4d_", table)); args[11] = par
    status, flux = integrate(*args)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "oracle"))
    import oracle as orc
    s2, ref = orc.integrate_time_invariance(*args)
    assert status == 0 and s2 == 0
    assert float(np.max(np.abs(flux - ref) / np.max(np.abs(ref)))) < PULSE_RTOL


def test_m4_hot_regions_with_elsewhere_correction_and_interstellar():
    """Config 4: hot regions with the elsewhere correction active (pyx:257-268,469-478), Elsewhere added to
    the first hot region (Photosphere.py:589-592), interstellar attenuation (Interstellar.py:27-58), fold and
    likelihood -- every stage on the GPU, against the reference's lnL."""
    import os
    from conftest import GOLDEN
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.cellmesh.integrator_for_azimuthal_invariance import integrate
    from xpsi_b200.cellmesh.integrator_for_time_invariance import integrate as integrate_tinv
    from xpsi_b200.instrument import fold
    from xpsi_b200.interstellar import attenuate
    from xpsi_b200.likelihoods import eval_marginal_likelihood
    from xpsi_b200.tools import energy_integrator
    d = np.load(os.path.join(GOLDEN, "m4_elsewhere.npz"))
    table = syn.nsx_like_table()
    matrix, edges = syn.nicer_like_response()[:2]
    status, spectrum = integrate_tinv(*_tinv_args(d, "else_", table))
    assert status == 0
    comps = []
    for m in range(2):
        p = "int%d_" % m
        args = list(_integrate_args(d, p, table))
        args[12] = d[p + "correction_srcCellParams"]
        args[23] = table
        args[25] = int(d[p + "else_atm_ext"])
        status, flux = integrate(*args)
        assert status == 0
        err = _pulse_err(flux, d[p + "flux"])
        print("M4 member", m, "flux (with correction) rel err", err)
        assert err < PULSE_RTOL
        if m == 0:
            flux = flux + spectrum[:, None]
        integrated = energy_integrator(1, flux / d["d_sq"], np.log10(d["int0_energies"]), np.log10(edges))
        attenuate(d["attenuation"], integrated)
        comps.append(fold(matrix, integrated, (0, matrix.shape[1]), (0, matrix.shape[0])))
        assert rel_err(comps[-1], d["marg_components_%d" % m]) < PULSE_RTOL
    p = "marg_"
    lnL = eval_marginal_likelihood(float(d[p + "exposure_time"]), d[p + "phases"], d[p + "counts"], tuple(comps),
                                   (d[p + "component_phases_0"], d[p + "component_phases_1"]), d[p + "phase_shifts"],
                                   d[p + "precomp"], d[p + "support"], 1000, 0.0, 1e-8, 1e-3, 10.0, -1e90)[0]
    print("M4 lnL", lnL, "ref", float(d["lnL_total"]), "diff", lnL - float(d["lnL_total"]))
    assert abs(lnL - float(d["lnL_total"])) < LNL_ATOL


def test_m3_cst_pdt_members_and_pipeline():
    """Config 3: CST primary (omission hole) + PDT secondary (superseding + ceding member): three integrator
    calls, members of a region summed after energy integration (Signal.py:419-429)."""
    import os
    from conftest import GOLDEN
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.cellmesh.integrator_for_azimuthal_invariance import integrate
    from xpsi_b200.pipeline import BatchedLikelihood
    d = np.load(os.path.join(GOLDEN, "m3_cst_pdt.npz"))
    table = syn.nsx_like_table()
    n_mem = int(d["n_members"])
    for m in range(n_mem):
        p = "int%d_" % m
        status, flux = integrate(*_integrate_args(d, p, table))
        assert status == 0
        err = _pulse_err(flux, d[p + "flux"])
        print("M3 member", m, "mesh", d[p + "cellArea"].shape, "flux rel err", err)
        assert err < PULSE_RTOL
    matrix, edges = syn.nicer_like_response()[:2]
    pipe = BatchedLikelihood(member_component=d["member_component"], max_rings=44, max_azi=44, n_rays=512,
                             energies=d["int0_energies"], leaves=d["int0_leaves"], phases=d["int0_phases"],
                             hot_atm_ext=2, hot_atmosphere=table, image_order_limit=3, response=matrix,
                             energy_edges=edges, counts=d["counts"], data_phases=np.linspace(0.0, 1.0, 33),
                             exposure_time=syn.M2_EXPOSURE, max_batch=4)
    B = 3
    batch = pipe.new_batch(B)
    for b in range(B):
        batch.omega[b] = d["int0_omega"]; batch.inclination[b] = d["int0_inclination"]; batch.d_sq[b] = d["d_sq"]
        batch.phase_shifts[b] = d["marg_phase_shifts"]
        for m in range(n_mem):
            g = lambda k: d["int%d_%s" % (m, k)]
            batch.set_member(b, m, g("cellArea"), g("theta"), g("phi"), g("radialCoords_of_parallels"),
                             g("r_s_over_r"), g("srcCellParams"), g("deflection"), g("cos_alpha"), g("lag"),
                             g("maxDeflection"), g("cos_gammaArray"))
    lnL, status = pipe(batch)
    print("M3 pipeline lnL", lnL, "ref", float(d["lnL_total"]), "status", status)
    assert (status == 0).all()
    assert np.max(np.abs(lnL - float(d["lnL_total"]))) < LNL_ATOL
    flux, folded, expected = pipe.fetch(B)
    for c in range(2):
        assert rel_err(folded[1, c], d["marg_components_%d" % c]) < PULSE_RTOL


def test_gpu_embed_matches_reference_mesh_and_rays(m2):
    """f1: mesh + rays of circular spots built on the GPU vs the reference's embed (golden fixture)."""
    from xpsi_b200 import synthetic as syn
    pipe = _m2_pipeline(m2, max_batch=4)
    thetas = np.array([m2["t0_theta"], m2["t1_theta"]])
    sb = syn.m2_spot_batch(pipe, thetas)
    pipe.embed_spots(sb)
    e = pipe.fetch_embed(2)
    for t in range(2):
        for m in range(2):
            q = t * 2 + m
            g = lambda k: m2["t%d_int%d_%s" % (t, m, k)]
            n = g("cellArea").shape[0]
            assert e["n_rings"][q] == n
            full = g("cellArea").max()
            errs = dict(
                theta=np.max(np.abs(e["theta"][q, :n] - g("theta")[:, 0])),
                phi=np.max(np.abs(e["phi"][q, :n, :n] - g("phi"))),
                radial=np.max(np.abs(e["radial"][q, :n] / g("radialCoords_of_parallels") - 1.0)),
                cos_gamma=np.max(np.abs(e["cos_gamma"][q, :n] - g("cos_gammaArray"))),
                params=np.max(np.abs(e["srcParams"][q, :n] - g("srcCellParams")[:, 0, :])),
                area=np.max(np.abs(e["cellArea"][q, :n, :n] - g("cellArea"))) / full,
                cos_alpha=np.max(np.abs(e["cos_alpha"][q, :n] - g("cos_alpha"))),
                deflection=np.max(np.abs(e["deflection"][q, :n] - g("deflection")) / np.maximum(g("deflection"), 1e-3)),
                lag=np.max(np.abs(e["lag"][q, :n] - g("lag"))) / np.max(np.abs(g("lag"))),
                maxDeflection=np.max(np.abs(e["maxDeflection"][q, :n] / g("maxDeflection") - 1.0)))
            print("embed theta", t, "member", m, {k: float("%.2e" % v) for k, v in errs.items()})
            assert ((e["cellArea"][q, :n, :n] > 0) == (g("cellArea") > 0)).all()
            assert errs["theta"] < 1e-12 and errs["phi"] < 1e-13 and errs["radial"] < 1e-13
            assert errs["cos_gamma"] < 1e-13 and errs["params"] < 1e-12
            assert errs["area"] < 1e-9          # the reference's own quadrature tolerance is 1e-8
            assert errs["cos_alpha"] < 1e-13 and errs["deflection"] < 1e-11 and errs["lag"] < 1e-10
            assert errs["maxDeflection"] < 1e-11


def test_theta_level_likelihood_with_gpu_embed(m2):
    """theta -> lnL entirely on the GPU (embed + integrate + fold + likelihood) vs the reference's lnL."""
    from xpsi_b200 import synthetic as syn
    pipe = _m2_pipeline(m2, max_batch=8)
    thetas = np.array([m2["t0_theta"], m2["t1_theta"], m2["t0_theta"]])
    lnL, status = pipe.eval_spots(syn.m2_spot_batch(pipe, thetas))
    assert (status == 0).all()
    for b, t in enumerate((0, 1, 0)):
        ref = float(m2["t%d_lnL_total" % t])
        print("theta-level lnL", lnL[b], "ref", ref, "diff", lnL[b] - ref)
        # measured 4.7e-9 (|lnL| = 3.8e4): the golden vectors' reference meshes are converged (tests/test_theta_parity.py
        # covers the prior draws whose reference mesh is not)
        assert abs(lnL[b] - ref) < THETA_GOLDEN_LNL_ATOL


def test_poisson_likelihood_given_background(c1):
    """row f4: Poisson likelihood with a given background + expected counts."""
    import os
    from conftest import GOLDEN
    from xpsi_b200.likelihoods import expected_counts, poisson_likelihood_given_background
    d = np.load(os.path.join(GOLDEN, "tools.npz"))
    comp = c1["marg_components_0"]
    lnL, expec = poisson_likelihood_given_background(1000.0, d["edges"], c1["marg_counts"], (comp,),
                                                     (d["sig_phases"],), np.array([float(d["plgb_shift"])]),
                                                     d["plgb_background"], c1["marg_precomp"])
    print("given-background lnL", lnL, "ref", float(d["plgb_lnL"]))
    assert abs(lnL - float(d["plgb_lnL"])) < LNL_ATOL
    assert rel_err(expec, d["plgb_expected"]) < PULSE_RTOL
    e2 = expected_counts(1000.0, d["edges"], (comp,), (d["sig_phases"],), np.array([float(d["plgb_shift"])]),
                         d["plgb_background"])
    assert rel_err(e2, d["plgb_expected"]) < PULSE_RTOL


def test_integrator_disc_beaming_and_steffen_options(c1, m2):
    """Optional branches of the integrator: disc occultation (common_functions.pyx:110-138), beaming options
    1-2 (hot_wrapper.pyx:155-172), Steffen phase interpolant (tools/core.pyx:34-52)."""
    from test_oracle import _option_cases
    from xpsi_b200 import tools
    from xpsi_b200.cellmesh.integrator_for_azimuthal_invariance import integrate
    cases, d = _option_cases(c1, m2)
    for name, a, kw, ref in cases:
        status, flux = integrate(*a, **kw)
        assert status == 0, name
        err = _pulse_err(flux, ref)
        print(name, "rel err", err)
        assert err < PULSE_RTOL, name
    try:
        tools.set_phase_interpolant("Steffen")
        status, flux = integrate(*_integrate_args(c1, "int0_", ()))
        err = _pulse_err(flux, d["steffen_c1"])
        print("Steffen phase interpolant rel err", err)
        assert status == 0 and err < PULSE_RTOL
    finally:
        tools.set_phase_interpolant("Akima")


def test_general_integrator_without_azimuthal_invariance(c1, m2):
    """cellmesh.integrator.integrate (integrator.pyx:48-667): blackbody, beaming, Num4D with uniform and with
    per-cell parameters, per-cell Num4D elsewhere correction, Steffen interpolant for the GEOM spline."""
    from test_oracle import _general_cases
    from xpsi_b200 import tools
    from xpsi_b200.cellmesh.integrator import integrate
    cases, d = _general_cases(c1, m2)
    for name, a, kw, ref in cases:
        status, flux = integrate(*a, **kw)
        assert status == 0, name
        err = _pulse_err(flux, ref)
        print("general:", name, "rel err", err)
        assert err < PULSE_RTOL, name
    try:
        tools.set_phase_interpolant("Steffen")
        status, flux = integrate(*cases[0][1])
        err = _pulse_err(flux, d["c1_steffen"])
        print("general: Steffen rel err", err)
        assert status == 0 and err < PULSE_RTOL
    finally:
        tools.set_phase_interpolant("Akima")
    # full energy grid (4 chunks of 32 energies) against the CPU oracle
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle as orc
    from xpsi_b200 import synthetic as syn
    a = list(_integrate_args(m2, "t1_int0_", syn.nsx_like_table()))
    s0, f0 = orc.integrate_general(*a)
    s1, f1 = integrate(*a)
    assert s0 == 0 and s1 == 0
    err = _pulse_err(f1, f0)
    print("general: m2 t1 member 0, 128 energies vs oracle", err)
    assert err < PULSE_RTOL


def test_surface_radiation_field_intensity():
    """surface_radiation_field.intensity (core.pyx:125-308): BB / Num4D, hot with beam_opt 0-3, elsewhere."""
    from test_oracle import _intensity_cases, _point_err
    from xpsi_b200.surface_radiation_field import intensity
    for name, a, ref in _intensity_cases():
        err = _point_err(intensity(*a), ref)
        print("intensity:", name, "point-wise rel err", err)
        assert err < PULSE_RTOL, name
    with pytest.raises(ValueError):
        intensity(np.ones(2), np.ones(2), np.ones((2, 2)), None, 0, 'nowhere', 'BB')
    with pytest.raises(ValueError):
        intensity(np.ones(2), np.ones(2), np.ones((2, 2)), None, 0, 'hot', 'Num4D')


def test_synthesise_expected_counts(c1):
    """tools.synthesise_exposure / synthesise_given_total_count_number (synthesise.pyx:40-276): the expected
    counts and normalisations against the reference; the Poisson draw is host-side and only sanity-checked."""
    from xpsi_b200 import tools
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tools.npz"))
    comp, sig_phases, edges = c1["marg_components_0"], c1["marg_component_phases_0"], c1["marg_phases"]
    e1, syn1, s1 = tools.synthesise_exposure(1000.0, edges, (comp,), (sig_phases,), np.array([0.13]), 5.0e4,
                                             d["syn_bg_counts"], gsl_seed=1)
    assert rel_err(e1, d["syn_exposure_expected"]) < PULSE_RTOL
    assert abs(s1 - float(d["syn_exposure_scale"])) <= 1e-12 * abs(s1)
    e2, syn2, a2, b2 = tools.synthesise_given_total_count_number(edges, 2.0e6, (comp,), (sig_phases,),
                                                                 np.array([0.13]), 5.0e4, d["syn_bg_counts"], gsl_seed=1)
    assert rel_err(e2, d["syn_total_expected"]) < PULSE_RTOL
    assert np.allclose([a2, b2], d["syn_total_scales"], rtol=1e-9, atol=0.0)
    assert abs(e2.sum() - 2.05e6) < 1e-3 * 2.05e6
    for syn, e in ((syn1, e1), (syn2, e2)):
        assert syn.shape == e.shape and np.all(syn >= 0) and np.all(syn == np.round(syn))
        assert abs(syn.sum() - e.sum()) < 6.0 * np.sqrt(e.sum())


def test_m4_batched_pipeline_with_elsewhere_and_interstellar(m2):
    """Config 4 through the batched pipeline: Elsewhere (time-invariant integrator + correction in the hot
    members + Photosphere.py:589-592), per-theta interstellar attenuation, fold, likelihood -- first with the
    reference's embedded arrays (mesh level), then from the parameter vector with every embed on the GPU."""
    from conftest import GOLDEN
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.pipeline import BatchedLikelihood
    d = np.load(os.path.join(GOLDEN, "m4_elsewhere.npz"))
    table = syn.nsx_like_table()
    matrix, edges = syn.nicer_like_response()[:2]
    pipe = BatchedLikelihood(member_component=[0, 1], max_rings=64, max_azi=64, n_rays=200,
                             energies=d["int0_energies"], leaves=d["int0_leaves"], phases=d["int0_phases"],
                             hot_atm_ext=2, hot_atmosphere=table, image_order_limit=3, response=matrix,
                             energy_edges=edges, counts=d["counts"], data_phases=np.linspace(0.0, 1.0, 33),
                             exposure_time=syn.M2_EXPOSURE, max_batch=4)
    N_H = float(d["column_density"])
    n = int(d["else_sqrt_numPix"])
    pipe.set_extras(elsewhere=dict(sqrt_num_cells=n, num_rays=int(d["else_numRays"]), atm_ext=2, atmosphere=table,
                                   image_order_limit=int(d["else_image_order_limit"])),
                    attenuation=d["attenuation"] ** (1.0 / N_H))
    B = 3
    ref = float(d["lnL_total"])
    # ---- mesh level -------------------------------------------------------------------------------
    batch = pipe.new_batch(B)
    corr = np.zeros((B * 2, 64, 2))
    for b in range(B):
        batch.omega[b] = d["int0_omega"]; batch.inclination[b] = d["int0_inclination"]; batch.d_sq[b] = d["d_sq"]
        batch.phase_shifts[b] = d["marg_phase_shifts"]
        for m in range(2):
            g = lambda k: d["int%d_%s" % (m, k)]
            batch.set_member(b, m, g("cellArea"), g("theta"), g("phi"), g("radialCoords_of_parallels"),
                             g("r_s_over_r"), g("srcCellParams"), g("deflection"), g("cos_alpha"), g("lag"),
                             g("maxDeflection"), g("cos_gammaArray"))
            c = g("correction_srcCellParams")
            J = np.argmax(g("cellArea") > 0.0, axis=1)
            corr[b * 2 + m, :c.shape[0]] = c[np.arange(c.shape[0]), J]
    tile = lambda a: np.ascontiguousarray(np.broadcast_to(a, (B,) + np.shape(a)))
    els = dict(cellArea=tile(float(d["else_cellArea"])), radial=tile(d["else_radialCoords_of_parallels"]),
               r_s_over_r=tile(d["else_r_s_over_r"]), theta=tile(d["else_theta"]), phi=tile(d["else_phi"]),
               srcParams=tile(d["else_srcCellParams"]), deflection=tile(d["else_deflection"]),
               cos_alpha=tile(d["else_cos_alpha"]), maxDeflection=tile(d["else_maxDeflection"]),
               cos_gamma=tile(d["else_cos_gammaArray"]))
    pipe.upload_extras(B, att_power=np.full(B, N_H), elsewhere=els, correction_srcParams=corr)
    lnL, status = pipe(batch)
    print("M4 pipeline (mesh level) lnL", lnL, "ref", ref, "status", status)
    assert (status == 0).all()
    assert np.max(np.abs(lnL - ref)) < LNL_ATOL
    spec = pipe.fetch_elsewhere(B)
    assert rel_err(spec[1], d["else_flux"]) < PULSE_RTOL
    flux, folded, expected = pipe.fetch(B)
    for c in range(2):
        assert rel_err(folded[2, c], d["marg_components_%d" % c]) < PULSE_RTOL
    # ---- parameter level: closed mesh, spots, rays and correction rows all embedded on the GPU -------
    thetas = np.tile(d["theta"], (B, 1))
    spots = syn.m2_spot_batch(pipe, thetas[:, :11])
    pipe.upload_extras(B, att_power=thetas[:, 12], else_temperature=thetas[:, 11])
    lnL2, status2 = pipe.eval_spots(spots)
    print("M4 pipeline (parameter level) lnL", lnL2, "diff", lnL2 - ref, "status", status2)
    assert (status2 == 0).all()
    assert np.max(np.abs(lnL2 - ref)) < THETA_GOLDEN_LNL_ATOL               # measured 2.6e-10
    spec2 = pipe.fetch_elsewhere(B)
    err = rel_err(spec2[0], d["else_flux"])
    print("GPU-embedded elsewhere spectrum rel err", err)
    assert err < PULSE_RTOL


def _embed_errors(e, q, g):
    """errors of embedded member q (pipeline.fetch_embed) against the reference's integrator inputs g(key)"""
    n = g("cellArea").shape[0]
    full = g("cellArea").max()
    return n, dict(
        theta=np.max(np.abs(e["theta"][q, :n] - g("theta")[:, 0])),
        phi=np.max(np.abs(e["phi"][q, :n, :n] - g("phi"))),
        radial=np.max(np.abs(e["radial"][q, :n] / g("radialCoords_of_parallels") - 1.0)),
        cos_gamma=np.max(np.abs(e["cos_gamma"][q, :n] - g("cos_gammaArray"))),
        params=np.max(np.abs(e["srcParams"][q, :n] - g("srcCellParams")[:, 0, :])),
        area=np.max(np.abs(e["cellArea"][q, :n, :n] - g("cellArea"))) / full,
        cos_alpha=np.max(np.abs(e["cos_alpha"][q, :n] - g("cos_alpha"))),
        deflection=np.max(np.abs(e["deflection"][q, :n] - g("deflection")) / np.maximum(g("deflection"), 1e-3)),
        lag=np.max(np.abs(e["lag"][q, :n] - g("lag"))) / np.max(np.abs(g("lag"))),
        maxDeflection=np.max(np.abs(e["maxDeflection"][q, :n] / g("maxDeflection") - 1.0)))


def _check_embed(e, q, g, tag):
    n, errs = _embed_errors(e, q, g)
    print("embed", tag, "mesh", n, {k: float("%.2e" % v) for k, v in errs.items()})
    assert e["n_rings"][q] == n, tag
    # the radiating set must agree except for slivers below the reference's own quadrature tolerance
    mism = (e["cellArea"][q, :n, :n] > 0) != (g("cellArea") > 0)
    if mism.any():
        print("   radiating-set mismatches:", int(mism.sum()), "largest area / full cell",
              float(np.maximum(e["cellArea"][q, :n, :n], g("cellArea"))[mism].max() / g("cellArea").max()))
    assert np.all(np.maximum(e["cellArea"][q, :n, :n], g("cellArea"))[mism] < 1e-5 * g("cellArea").max()), tag
    assert errs["theta"] < 1e-12 and errs["phi"] < 1e-13 and errs["radial"] < 1e-13, tag
    assert errs["cos_gamma"] < 1e-13 and errs["params"] < 1e-12, tag
    both = (e["cellArea"][q, :n, :n] > 0) & (g("cellArea") > 0)
    err_both = np.max(np.abs(e["cellArea"][q, :n, :n] - g("cellArea"))[both]) / g("cellArea").max()
    print("   area error over cells radiating in both", float(err_both))
    assert err_both < 1e-7, tag                # the reference integrates cell areas with CQUAD at epsrel 1e-8
    assert errs["area"] < 1e-5, tag            # slivers the reference's quadrature misses entirely
    assert errs["cos_alpha"] < 1e-13 and errs["deflection"] < 1e-11 and errs["lag"] < 1e-10, tag
    assert errs["maxDeflection"] < 1e-11, tag


def test_gpu_embed_of_omission_ceding_and_polar_regions(m2):
    """f1: CST + PDT hot regions (omission hole; superseding + ceding members sharing the cell budget) and polar
    caps (polar_mesh.pyx) embedded on the GPU from parameter values, against the reference's embed and lnL."""
    from conftest import GOLDEN
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.pipeline import BatchedLikelihood
    table = syn.nsx_like_table()
    matrix, edges = syn.nicer_like_response()[:2]
    # ---- M3: primary CST, secondary PDT -----------------------------------------------------------------
    d = np.load(os.path.join(GOLDEN, "m3_cst_pdt.npz"))
    th = dict(zip([str(n) for n in d["names"]], d["theta"]))
    pipe = BatchedLikelihood(member_component=[0, 1, 1], max_rings=80, max_azi=80, n_rays=512,
                             energies=d["int0_energies"], leaves=d["int0_leaves"], phases=d["int0_phases"],
                             hot_atm_ext=2, hot_atmosphere=table, image_order_limit=3, response=matrix,
                             energy_edges=edges, counts=d["counts"], data_phases=np.linspace(0.0, 1.0, 33),
                             exposure_time=syn.M2_EXPOSURE, max_batch=2)
    B = 2
    sb = pipe.new_spot_batch(B, syn.M2_FREQUENCY, num_cells=1024, min_sqrt_num_cells=10, max_sqrt_num_cells=80)
    one = np.ones(B)
    sb.set_spacetime(th["mass"] * one, th["radius"] * one, th["distance"] * one, th["cos_inclination"] * one,
                     syn.M2_FREQUENCY)
    sb.phase_shifts[:, 0], sb.phase_shifts[:, 1] = th["p__phase_shift"], th["s__phase_shift"]
    sb.set_region(0, super_colatitude=th["p__super_colatitude"] * one, super_radius=th["p__super_radius"] * one,
                  super_temperature=th["p__super_temperature"] * one, omit_radius=th["p__omit_radius"] * one)
    sb.set_region(1, 2, super_colatitude=th["s__super_colatitude"] * one, super_radius=th["s__super_radius"] * one,
                  super_temperature=th["s__super_temperature"] * one, cede_colatitude=th["s__cede_colatitude"] * one,
                  cede_radius=th["s__cede_radius"] * one, cede_azimuth=th["s__cede_azimuth"] * one,
                  cede_temperature=th["s__cede_temperature"] * one, is_antiphased=True)
    lnL, status = pipe.eval_spots(sb)
    e = pipe.fetch_embed(B)
    for m in range(3):
        _check_embed(e, 3 + m, lambda k: d["int%d_%s" % (m, k)], "M3 member %d" % m)
    print("M3 parameter-level lnL", lnL, "ref", float(d["lnL_total"]), "diff", lnL - float(d["lnL_total"]), "status", status)
    assert (status == 0).all()
    assert np.max(np.abs(lnL - float(d["lnL_total"]))) < LNL_ATOL
    # ---- polar caps ---------------------------------------------------------------------------------------
    d = np.load(os.path.join(GOLDEN, "m5_polar.npz"))
    pipe = BatchedLikelihood(member_component=[0, 1], max_rings=64, max_azi=64, n_rays=200,
                             energies=d["int0_energies"], leaves=d["int0_leaves"], phases=d["int0_phases"],
                             hot_atm_ext=2, hot_atmosphere=table, image_order_limit=3, response=matrix,
                             energy_edges=edges, counts=d["counts"], data_phases=np.linspace(0.0, 1.0, 33),
                             exposure_time=syn.M2_EXPOSURE, max_batch=2)
    thetas = np.tile(d["theta"], (2, 1))
    lnL, status = pipe.eval_spots(syn.m2_spot_batch(pipe, thetas))
    e = pipe.fetch_embed(2)
    for m in range(2):
        _check_embed(e, m, lambda k: d["int%d_%s" % (m, k)], "polar member %d" % m)
    print("polar-cap parameter-level lnL", lnL, "ref", float(d["lnL_total"]), "diff", lnL - float(d["lnL_total"]), "status", status)
    assert (status == 0).all()
    assert np.max(np.abs(lnL - float(d["lnL_total"]))) < LNL_ATOL


def test_likelihood_callable_mirrors_reference_conventions(m2):
    """xpsi.Likelihood.__call__ (Likelihood.py:450-511) over the GPU pipeline: value, prior handling,
    random-near-llzero on rejected / failed points, batched evaluation."""
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.likelihood import Likelihood
    pipe = _m2_pipeline(m2, max_batch=8)
    fill = lambda pl, P: syn.m2_spot_batch(pl, P)
    like = Likelihood(pipe, fill, llzero=-1.0e90)
    ref = float(m2["t0_lnL_total"])
    v = like(m2["t0_theta"], force=True)
    assert isinstance(v, float) and abs(v - ref) < THETA_GOLDEN_LNL_ATOL
    assert like(m2["t0_theta"]) == v                                  # memoised
    assert like() == v                                                # externally updated / cached vector
    with_prior = Likelihood(pipe, fill, prior=lambda p: -3.5)
    assert abs(with_prior(m2["t1_theta"]) - (float(m2["t1_lnL_total"]) - 3.5)) < THETA_GOLDEN_LNL_ATOL
    rejected = Likelihood(pipe, fill, prior=lambda p: -np.inf)
    r = rejected(m2["t0_theta"])
    assert -1.0e90 <= r <= -1.0e89                                    # Likelihood.py:267-271
    lnL, status = like.batch(np.array([m2["t0_theta"], m2["t1_theta"], m2["t0_theta"]]))
    assert (status == 0).all() and abs(lnL[1] - float(m2["t1_lnL_total"])) < THETA_GOLDEN_LNL_ATOL and abs(lnL[0] - lnL[2]) < 1e-7
    with pytest.raises(TypeError):
        Likelihood(pipe, fill)()


def test_edge_cases_and_error_paths(c1):
    """Edge cases the reference handles or rejects: meshes without radiating cells, dark rings, one image order,
    a single energy; argument errors surface as exceptions with the library's message (never a silent fallback)."""
    from xpsi_b200 import _lib
    from xpsi_b200.cellmesh.integrator_for_azimuthal_invariance import integrate
    from xpsi_b200.cellmesh.integrator import integrate as integrate_general
    from xpsi_b200.tools import energy_integrator
    base = list(_integrate_args(c1, "int0_", ()))
    # nothing radiates -> zeros, status 0 (pyx:286-296 skips every ring)
    a = list(base); a[11] = np.zeros_like(base[11])
    for fn in (integrate, integrate_general):
        status, flux = fn(*a)
        assert status == 0 and flux.shape == c1["int0_flux"].shape and not flux.any()
    # only a few rings radiate -> equals the sum restricted to those rings (rings are independent)
    rad = np.zeros_like(base[11]); rad[3:6] = base[11][3:6]
    a = list(base); a[11] = rad
    s1, f_part = integrate(*a)
    rad2 = np.array(base[11]); rad2[3:6] = 0
    a = list(base); a[11] = rad2
    s2, f_rest = integrate(*a)
    s3, f_all = integrate(*base)
    assert s1 == 0 and s2 == 0 and s3 == 0
    assert _pulse_err(f_part + f_rest, f_all) < 1e-13
    # image_order_limit = 1 keeps the primary image only and can only lower the flux
    a = list(base); a[27] = 1
    s4, f_one = integrate(*a)
    assert s4 == 0 and np.all(f_one <= f_all * (1 + 1e-12)) and f_one.sum() > 0.9 * f_all.sum()
    # a single energy is a valid call
    a = list(base); a[19] = np.ascontiguousarray(base[19][40:41])
    s5, f_single = integrate(*a)
    assert s5 == 0 and _pulse_err(f_single, f_all[40:41]) < 1e-13
    # argument errors
    a = list(base); a[20] = np.ascontiguousarray(base[20][:4])        # fewer than 5 leaves
    with pytest.raises(_lib.XpsiB200Error):
        integrate(*a)
    a = list(base); a[24] = 7                                          # unknown atmosphere extension
    with pytest.raises(NotImplementedError):
        integrate(*a)
    a = list(base); a[21] = np.linspace(0.0, 2 * np.pi, 200)           # more phases than the kernels cover
    with pytest.raises((_lib.XpsiB200Error, NotImplementedError)):
        integrate(*a)
    with pytest.raises(_lib.XpsiB200Error):                            # Akima needs >= 5 energies
        energy_integrator(1, np.ones((3, 4)), np.log10([1.0, 2.0, 3.0]), np.log10([1.0, 1.5, 2.5]))


def test_pipeline_properties_at_bench_batch_size(m2):
    """Size-independent properties on a bench-sized batch (512 parameter vectors): results do not depend on the
    position in the batch, a whole-cycle phase shift changes nothing, and the flux scales as 1/d^2."""
    from xpsi_b200 import synthetic as syn
    pipe = _m2_pipeline(m2, max_batch=512)
    B = 512
    thetas = np.vstack([m2["t0_theta"], m2["t1_theta"], syn.m2_theta_batch(B - 2)])
    lnL, status = pipe.eval_spots(syn.m2_spot_batch(pipe, thetas))
    assert np.isin(status, (0, 11)).all()
    assert abs(lnL[0] - float(m2["t0_lnL_total"])) < THETA_GOLDEN_LNL_ATOL and abs(lnL[1] - float(m2["t1_lnL_total"])) < THETA_GOLDEN_LNL_ATOL
    perm = np.random.default_rng(0).permutation(B)
    lnL_p, status_p = pipe.eval_spots(syn.m2_spot_batch(pipe, thetas[perm]))
    ok = (status == 0)
    assert (status_p == status[perm]).all()
    assert np.max(np.abs(lnL_p[ok[perm]] - lnL[perm][ok[perm]])) < 1e-6      # ring sums use fp64 atomics (DESIGN s5.4)
    shifted = np.array(thetas[:64]); shifted[:, 4] += 1.0; shifted[:, 8] -= 1.0
    lnL_s, status_s = pipe.eval_spots(syn.m2_spot_batch(pipe, shifted))
    good = (status[:64] == 0) & (status_s == 0)
    assert good.sum() > 10 and np.max(np.abs(lnL_s[good] - lnL[:64][good])) < 1e-6
    flux_a = pipe.fetch(64, folded=False, expected=False)[0]
    far = np.array(thetas[:64]); far[:, 2] *= 2.0
    pipe.eval_spots(syn.m2_spot_batch(pipe, far))
    folded_far = pipe.fetch(64, flux=False, expected=False)[1]
    pipe.eval_spots(syn.m2_spot_batch(pipe, thetas[:64]))
    flux_b, folded_near, _ = pipe.fetch(64, expected=False)
    assert rel_err(folded_far * 4.0, folded_near) < 1e-12                       # Likelihood.py:361-364
    assert rel_err(flux_b, flux_a) < 1e-10


def test_marginal_likelihood_with_more_than_32_phase_bins(c1):
    """eval_marginal_likelihood with 64 and with a ragged 45 data phase bins (lanes own several bins)."""
    from xpsi_b200.likelihoods import eval_marginal_likelihood, precomputation
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tools.npz"))
    comp, sig_phases = c1["marg_components_0"], c1["marg_component_phases_0"]
    sup = -1.0 * np.ones((comp.shape[0], 2)); sup[:, 0] = 0.0
    for nb in (64, 45):
        ph = np.linspace(0.0, 1.0, nb + 1)
        cnts = d["mbins%d_counts" % nb]
        pre = precomputation(cnts.astype(np.int32))
        res = eval_marginal_likelihood(1000.0, ph, cnts, (comp,), (sig_phases,), np.array([0.13]), pre, sup,
                                       1000, 0.0, 1.0e-8, 1.0e-3, 10.0, -1.0e90)
        print(nb, "bins: lnL", res[0], "ref", float(d["mbins%d_lnL" % nb]))
        assert abs(res[0] - float(d["mbins%d_lnL" % nb])) < LNL_ATOL
        assert rel_err(res[1], d["mbins%d_expected" % nb]) < PULSE_RTOL
        assert rel_err(res[2], d["mbins%d_bg" % nb]) < 1e-6


def test_odd_leaf_counts_and_unequal_phase_grids(c1):
    """N_L odd (the middle leaf is its own mirror, pyx:416-419), N_P != N_L up to the 128-phase limit."""
    from test_oracle import _odd_leaf_cases
    from xpsi_b200.cellmesh.integrator_for_azimuthal_invariance import integrate
    from xpsi_b200.cellmesh.integrator import integrate as integrate_general
    for nl, nph, a, ref_az, ref_gen in _odd_leaf_cases(c1):
        status, flux = integrate(*a)
        err = _pulse_err(flux, ref_az)
        print("N_L", nl, "N_P", nph, "azimuthal-invariance rel err", err)
        assert status == 0 and err < PULSE_RTOL
        status, flux = integrate_general(*a)
        err = _pulse_err(flux, ref_gen)
        print("N_L", nl, "N_P", nph, "general rel err", err)
        assert status == 0 and err < PULSE_RTOL


def test_known_answer_from_parameters_entirely_on_gpu(c1):
    """The reference's published known answer (xpsi/tests/test_likelihood.py:134: lnL = -47881.27817666349 for
    theta = [1.4, 10, 1, cos 60deg, 0, 70deg, 0.75, 6.8], blackbody ST model, 314 Hz) reproduced from the
    parameter vector with every stage -- embed, integrate, fold, likelihood -- on the GPU."""
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.pipeline import BatchedLikelihood
    matrix, edges = syn.c1_response()[:2]
    d, p = c1, "marg_"
    pipe = BatchedLikelihood(member_component=[0], max_rings=64, max_azi=64, n_rays=int(c1["int0_numRays"]),
                             energies=c1["int0_energies"], leaves=c1["int0_leaves"], phases=c1["int0_phases"],
                             hot_atm_ext=1, image_order_limit=int(c1["int0_image_order_limit"]), response=matrix,
                             energy_edges=edges, counts=d[p + "counts"], data_phases=d[p + "phases"],
                             exposure_time=float(d[p + "exposure_time"]), support=d[p + "support"], max_batch=4)
    th = c1["theta"]
    B = 3
    sb = pipe.new_spot_batch(B, 314.0, num_cells=1024, min_sqrt_num_cells=16, max_sqrt_num_cells=64)
    one = np.ones(B)
    sb.set_spacetime(th[0] * one, th[1] * one, th[2] * one, th[3] * one, 314.0)
    sb.phase_shifts[:, 0] = th[4]
    sb.colatitude[:, 0], sb.ang_radius[:, 0], sb.temperature[:, 0] = th[5], th[6], th[7]
    sb.phi_shift[:, 0] = np.pi                                         # is_antiphased (main.py:122-128)
    lnL, status = pipe.eval_spots(sb)
    known = float(c1["known_answer_lnL"])
    print("known-answer lnL from theta", lnL, "reference", known, "diff", lnL - known)
    assert (status == 0).all()
    assert known == -47881.27817666349
    assert np.max(np.abs(lnL - known)) < 1e-5 * abs(known)            # the reference's own rtol
    assert np.max(np.abs(lnL - float(c1["lnL_total"]))) < LNL_ATOL    # and the 1e-6 bar against the shim build
    e = pipe.fetch_embed(B)
    _check_embed(e, 1, lambda k: c1["int0_" + k], "C1 spot")


def test_theta_level_pipeline_with_beaming_parameters(m2):
    """Custom hot regions with beaming parameters (CustomHotRegion_Beaming.py:149-178: srcCellParams columns
    T, g, abb, bbb, cbb, dbb, nimu) through the parameter-level pipeline; the secondary's flux is checked against
    the reference integrator's output for the same beaming parameters (options.npz)."""
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.pipeline import BatchedLikelihood
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "options.npz"))
    matrix, edges = syn.nicer_like_response()[:2]
    beam = d["beam_params_m2"][0, 0, 2:]
    for opt in (2, 3):
        pipe = BatchedLikelihood(member_component=[0, 1], max_rings=64, max_azi=64, n_rays=200,
                                 energies=m2["t0_int0_energies"], leaves=m2["t0_int0_leaves"],
                                 phases=m2["t0_int0_phases"], hot_atm_ext=2, hot_atmosphere=syn.nsx_like_table(),
                                 image_order_limit=3, response=matrix, energy_edges=edges, counts=m2["counts"],
                                 data_phases=np.linspace(0.0, 1.0, 33), exposure_time=syn.M2_EXPOSURE,
                                 n_params=7, max_batch=2)
        pipe.set_extras(beam_opt=opt)
        spots = syn.m2_spot_batch(pipe, np.array([m2["t0_theta"], m2["t0_theta"]]))
        spots.extra_params = np.tile(beam, (2, 2, 1))
        lnL, status = pipe.eval_spots(spots)
        assert (status == 0).all() or np.isin(status, (0, 11)).all()
        flux = pipe.fetch(2, folded=False, expected=False)[0]
        got = flux[3] / (m2["t0_int0_energies"][:, None] * 1.60217662e-16)      # theta 1, member 1
        err = _pulse_err(got, d["beam%d_m2" % opt])
        print("beam_opt", opt, "parameter-level flux vs reference integrator rel err", err, "lnL", lnL)
        assert err < PULSE_RTOL
        assert abs(lnL[0] - lnL[1]) < 1e-6


def test_cubic_interpolants(c1):
    """'Cubic' interpolants (tools/core.pyx:84-116: cspline_periodic in phase, natural cspline in energy) in every
    kernel that builds a spline: both pulse integrators, energy integrator, phase / energy tools, likelihood."""
    from xpsi_b200 import tools
    from xpsi_b200.cellmesh.integrator_for_azimuthal_invariance import integrate
    from xpsi_b200.cellmesh.integrator import integrate as integrate_general
    from xpsi_b200.likelihoods import eval_marginal_likelihood
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "options.npz"))
    a = list(_integrate_args(c1, "int0_", ()))
    a[19] = np.ascontiguousarray(c1["int0_energies"][::8])
    comp = c1["marg_components_0"]
    pulse = np.ascontiguousarray(comp[::8])
    try:
        tools.set_phase_interpolant("Cubic")
        tools.set_energy_interpolant("Cubic")
        s, f = integrate(*a)
        err = _pulse_err(f, d["cubic_azinv"]); print("Cubic azimuthal-invariance integrator", err)
        assert s == 0 and err < PULSE_RTOL
        s, f = integrate_general(*a)
        err = _pulse_err(f, d["cubic_general"]); print("Cubic general integrator", err)
        assert s == 0 and err < PULSE_RTOL
        sig = np.ascontiguousarray(c1["int0_flux"] / c1["d_sq"])
        e = rel_err(tools.energy_integrator(1, sig, c1["eint_log10_energies"], c1["eint_log10_edges"]), d["cubic_eint"])
        print("Cubic energy integrator", e); assert e < PULSE_RTOL
        e = rel_err(tools.phase_integrator(1000.0, c1["marg_phases"], pulse, c1["marg_component_phases_0"], 0.37),
                    d["cubic_pint"])
        print("Cubic phase integrator", e); assert e < PULSE_RTOL
        e = rel_err(tools.phase_interpolator(np.linspace(0.0, 1.0, 41), c1["marg_component_phases_0"], pulse, -0.2),
                    d["cubic_pitp"])
        print("Cubic phase interpolator", e); assert e < PULSE_RTOL
        flux4 = np.ascontiguousarray(c1["int0_flux"][:, ::4])
        e = rel_err(tools.energy_interpolator(1, flux4, np.log10(c1["int0_energies"]), d["cubic_new_E"]), d["cubic_eitp"])
        print("Cubic energy interpolator", e); assert e < PULSE_RTOL
        res = eval_marginal_likelihood(float(c1["marg_exposure_time"]), c1["marg_phases"], c1["marg_counts"], (comp,),
                                       (c1["marg_component_phases_0"],), c1["marg_phase_shifts"], c1["marg_precomp"],
                                       c1["marg_support"], 1000, 0.0, 1.0e-8, 1.0e-3, 10.0, -1.0e90)
        print("Cubic lnL", res[0], "ref", float(d["cubic_lnL"]))
        assert abs(res[0] - float(d["cubic_lnL"])) < LNL_ATOL
    finally:
        tools.set_phase_interpolant("Akima")
        tools.set_energy_interpolant("Steffen")


def test_time_invariant_component_likelihood():
    """Everywhere(time_invariant=True): a single signal column with phase-averaged data -- the rate is stored in the
    first bin without a phase spline (compute_expected_counts.pyx:190-192)."""
    from xpsi_b200.likelihoods import eval_marginal_likelihood, precomputation
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tools.npz"))
    folded, cnts = d["tinv_folded"], d["tinv_counts"]
    sup = -1.0 * np.ones((folded.shape[0], 2)); sup[:, 0] = 0.0
    res = eval_marginal_likelihood(float(d["tinv_exposure"]), np.array([0.0, 1.0]), cnts, (folded,), (np.array([0.0]),),
                                   np.array([0.0]), precomputation(cnts.astype(np.int32)), sup,
                                   1000, 0.0, 1.0e-8, 1.0e-3, 10.0, -1.0e90)
    print("time-invariant lnL", res[0], "ref", float(d["tinv_lnL"]))
    assert abs(res[0] - float(d["tinv_lnL"])) < LNL_ATOL
    assert rel_err(res[1], d["tinv_expected"]) < PULSE_RTOL and rel_err(res[2], d["tinv_bg"]) < 1e-6


@pytest.mark.parametrize("shape,stride", [((5, 4, 9, 41), 1), ((6, 5, 67, 165), 1), ((4, 4, 5, 23), 1),
                                          ((35, 14, 67, 166), 4), ((7, 4, 33, 91), 3)])
def test_num4d_tables_of_odd_and_small_shapes(m2, shape, stride):
    """Atmosphere tables whose axes are odd-sized, short or coarse: the slab tile a flux CTA fetches (one TMA box
    of [n_mu][rows] doubles starting on an even row) then runs past the ring's slab row or past the end of the
    table, and the row budgets shrink to a handful of rows.  Checked against the oracle on the M2 mesh and rays."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle as orc
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.cellmesh.integrator_for_azimuthal_invariance import integrate
    table = syn.nsx_like_table(shape)
    args = list(_integrate_args(m2, "t0_int0_", table))
    args[19] = np.ascontiguousarray(args[19][::stride])        # sparser energies: wider chunks, more rows per tile
    status, flux = integrate(*args)
    ref_status, ref = orc.integrate(*args)
    assert status == 0 and ref_status == 0
    err = _pulse_err(flux, ref)
    print("table shape", shape, "energy stride", stride, "flux rel err vs oracle", err)
    assert err < PULSE_RTOL
