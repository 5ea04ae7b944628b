"""Pin the CPU oracle (oracle/xpsi_oracle.c on the GSL-subset shim) against the
golden vectors recorded from the reference's own sources (tests/golden/)."""
import os
import sys

import numpy as np

from conftest import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402


def _integrate_args(d, prefix, atmosphere):
    g = lambda k: d[prefix + k]
    return (1, float(g("R")), float(g("omega")), float(g("r_s")), float(g("inclination")),
            g("cellArea"), g("radialCoords_of_parallels"), g("r_s_over_r"), g("theta"), g("phi"),
            g("srcCellParams"), g("CELL_RADIATES"), None, int(g("numRays")), g("deflection"),
            g("cos_alpha"), g("lag"), g("maxDeflection"), g("cos_gammaArray"), g("energies"),
            g("leaves"), g("phases"), atmosphere, (), int(g("hot_atm_ext")), 1, int(g("beam_opt")),
            int(g("image_order_limit")))


def _pulse_err(out, ref):
    scale = np.max(np.abs(ref), axis=1, keepdims=True)
    scale[scale == 0.0] = 1.0
    return float(np.max(np.abs(out - ref) / scale))


def test_oracle_integrate_blackbody_c1(c1):
    status, flux = orc.integrate(*_integrate_args(c1, "int0_", ()))
    assert status == 0
    assert _pulse_err(flux, c1["int0_flux"]) < 1e-12


def test_oracle_integrate_num4d_m2(m2):
    from xpsi_b200 import synthetic as syn
    table = syn.nsx_like_table()
    for t, m in ((0, 0), (0, 1), (1, 0)):
        prefix = "t%d_int%d_" % (t, m)
        status, flux = orc.integrate(*_integrate_args(m2, prefix, table))
        assert status == 0
        # stateless stencil vs the reference's history-dependent one: see xpsi_oracle.c header
        assert _pulse_err(flux, m2[prefix + "flux"]) < 1e-8


def test_oracle_energy_integrator_and_fold(c1, m2):
    from xpsi_b200 import synthetic as syn
    out = orc.energy_integrator(1, c1["int0_flux"] / c1["d_sq"], c1["eint_log10_energies"], c1["eint_log10_edges"])
    assert rel_err(out, c1["eint0_out"]) < 1e-13
    folded = orc.fold(syn.c1_response()[0], c1["eint0_out"])
    assert rel_err(folded, c1["marg_components_0"]) < 1e-13
    out = orc.energy_integrator(1, m2["t0_int0_flux"] / m2["t0_d_sq"], m2["t0_eint_log10_energies"],
                                m2["t0_eint_log10_edges"])
    assert rel_err(out, m2["t0_eint0_out"]) < 1e-13


def _marginal(d, p):
    n = int(d[p + "n_components"])
    comps = tuple(d["%scomponents_%d" % (p, i)] for i in range(n))
    cph = tuple(d["%scomponent_phases_%d" % (p, i)] for i in range(n))
    return orc.eval_marginal_likelihood(float(d[p + "exposure_time"]), d[p + "phases"], d[p + "counts"], comps,
                                        cph, d[p + "phase_shifts"], d[p + "precomp"], d[p + "support"], 1000,
                                        0.0, 1e-8, 1e-3, 10.0, -1e90)


def test_oracle_marginal_likelihood(c1, m2):
    rc, lnL, star, mcl, mcls = _marginal(c1, "marg_")
    assert rc == 0
    assert abs(lnL - float(c1["marg_lnL"])) < 1e-9
    assert abs(lnL + 47881.27817666349) < 1e-5 * 47881.27817666349     # published known answer
    assert rel_err(star, c1["marg_expected_counts"]) < 1e-12
    for t in range(int(m2["n_theta"])):
        rc, lnL, star, mcl, mcls = _marginal(m2, "t%d_marg_" % t)
        assert rc == 0 and abs(lnL - float(m2["t%d_marg_lnL" % t])) < 1e-8


def test_oracle_precomputation(c1):
    out = orc.precomputation(c1["marg_counts"].astype(np.int32))
    assert np.max(np.abs(out - c1["marg_precomp"])) < 1e-9


def test_gslshim_against_scipy():
    """Akima / natural + periodic cubic splines of the shim vs scipy's."""
    import ctypes as C
    from scipy.interpolate import Akima1DInterpolator, CubicSpline
    L = orc.lib
    dp = orc.dp
    L.gsl_interp_alloc.restype = C.c_void_p
    L.gsl_interp_alloc.argtypes = [C.c_void_p, C.c_size_t]
    L.gsl_interp_init.argtypes = [C.c_void_p, dp, dp, C.c_size_t]
    L.gsl_interp_eval.restype = C.c_double
    L.gsl_interp_eval.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_void_p]
    L.gsl_interp_eval_integ.restype = C.c_double
    L.gsl_interp_eval_integ.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_double, C.c_void_p]
    rng = np.random.default_rng(0)
    x = np.sort(rng.uniform(0, 10, 40))
    y = np.sin(x) + 0.1 * rng.normal(size=40)
    q = rng.uniform(x[0], x[-1], 100)
    xp, yp = x.ctypes.data_as(dp), y.ctypes.data_as(dp)

    def mk(name, yy):
        it = L.gsl_interp_alloc(C.c_void_p.in_dll(L, name), len(x))
        L.gsl_interp_init(it, xp, yy.ctypes.data_as(dp), len(x))
        return it
    ak = mk("gsl_interp_akima", y)
    ref = Akima1DInterpolator(x, y)
    assert max(abs(L.gsl_interp_eval(ak, xp, yp, v, None) - ref(v)) for v in q) < 1e-13
    assert abs(L.gsl_interp_eval_integ(ak, xp, yp, 1.0, 9.0, None) - ref.integrate(1.0, 9.0)) < 1e-12
    cs = mk("gsl_interp_cspline", y)
    ref = CubicSpline(x, y, bc_type="natural")
    assert max(abs(L.gsl_interp_eval(cs, xp, yp, v, None) - ref(v)) for v in q) < 1e-12
    y2 = y.copy(); y2[-1] = y2[0]
    y2p = y2.ctypes.data_as(dp)
    cp = mk("gsl_interp_cspline_periodic", y2)
    ref = CubicSpline(x, y2, bc_type="periodic")
    assert max(abs(L.gsl_interp_eval(cp, xp, y2p, v, None) - ref(v)) for v in q) < 1e-12


def _tag(shift):
    return ("%+.2f" % shift).replace(".", "p").replace("+", "P").replace("-", "M")


def test_oracle_signal_tools():
    d = np.load(os.path.join(ROOT, "tests", "golden", "tools.npz"))
    for shift in d["shifts"]:
        out = orc.phase_integrator(1000.0, d["edges"], d["pulse"], d["sig_phases"], float(shift))
        assert rel_err(out, d["pint_" + _tag(shift)]) < 1e-13
        out = orc.phase_interpolator(d["new_phases"], d["sig_phases"], d["pulse"], float(shift))
        assert rel_err(out, d["pitp_" + _tag(shift)]) < 1e-13
    assert rel_err(orc.energy_interpolator(1, d["flux"], d["log10E"], d["new_E"]), d["eitp"]) < 1e-12
    assert rel_err(orc.energy_interpolator(1, d["flux_neg"], d["log10E"], d["new_E"]), d["eitp_neg"]) < 1e-12


def _tinv_args(d, prefix, atmosphere):
    g = lambda k: d[prefix + k]
    return (1, float(g("R")), float(g("omega")), float(g("r_s")), float(g("inclination")), int(g("sqrt_numPix")),
            float(g("cellArea")), g("radialCoords_of_parallels"), g("r_s_over_r"), g("theta"), g("phi"),
            g("srcCellParams"), int(g("numRays")), g("deflection"), g("cos_alpha"), g("maxDeflection"),
            g("cos_gammaArray"), g("energies"), atmosphere, int(g("atm_ext")), int(g("image_order_limit")))


def test_oracle_time_invariant_integrator():
    from xpsi_b200 import synthetic as syn
    table = syn.nsx_like_table()
    ev = np.load(os.path.join(ROOT, "tests", "golden", "m4_everywhere.npz"))
    s, f = orc.integrate_time_invariance(*_tinv_args(ev, "bb_", ()))
    assert s == 0 and rel_err(f, ev["bb_flux"]) < 1e-12
    s, f = orc.integrate_time_invariance(*_tinv_args(ev, "num4d_", table))
    assert s == 0 and rel_err(f, ev["num4d_flux"]) < 1e-8
    el = np.load(os.path.join(ROOT, "tests", "golden", "m4_elsewhere.npz"))
    s, f = orc.integrate_time_invariance(*_tinv_args(el, "else_", table))
    assert s == 0 and rel_err(f, el["else_flux"]) < 1e-8


def test_oracle_integrate_with_elsewhere_correction():
    from xpsi_b200 import synthetic as syn
    table = syn.nsx_like_table()
    d = np.load(os.path.join(ROOT, "tests", "golden", "m4_elsewhere.npz"))
    for m in range(2):
        p = "int%d_" % m
        args = list(_integrate_args(d, p, table))
        args[12] = d[p + "correction_srcCellParams"]; args[23] = table; args[25] = int(d[p + "else_atm_ext"])
        status, flux = orc.integrate(*args)
        assert status == 0
        assert _pulse_err(flux, d[p + "flux"]) < 1e-8


def _option_cases(c1, m2):
    from xpsi_b200 import synthetic as syn
    d = np.load(os.path.join(ROOT, "tests", "golden", "options.npz"))
    table = syn.nsx_like_table()
    base = list(_integrate_args(c1, "int0_", ()))
    cases = []
    for R_in in (2.0e4, 5.0e4):
        a = list(base); a[8] = d["disk_theta"]
        cases.append(("disc R_in=%g" % R_in, a, dict(R_in=R_in), d["disk_flux_%d" % int(R_in)]))
    for opt in (1, 2, 3):
        a = list(base); a[10] = d["beam_params_c1"]; a[26] = opt
        cases.append(("beam %d BB" % opt, a, {}, d["beam%d_c1" % opt]))
        a = list(_integrate_args(m2, "t0_int1_", table)); a[10] = d["beam_params_m2"]; a[26] = opt
        cases.append(("beam %d Num4D" % opt, a, {}, d["beam%d_m2" % opt]))
    return cases, d


def test_oracle_disc_and_beaming_options(c1, m2):
    cases, d = _option_cases(c1, m2)
    for name, a, kw, ref in cases:
        status, flux = orc.integrate(*a, **kw)
        assert status == 0, name
        assert _pulse_err(flux, ref) < 1e-8, name
    status, flux = orc.integrate(*_integrate_args(c1, "int0_", ()), phase_interpolant="Steffen")
    assert status == 0 and _pulse_err(flux, d["steffen_c1"]) < 1e-12


def _general_cases(c1, m2):
    """inputs + reference outputs of tests/golden/make_golden_general.py (integrator.pyx, energies thinned)"""
    from xpsi_b200 import synthetic as syn
    d = np.load(os.path.join(ROOT, "tests", "golden", "general.npz"))
    m4 = np.load(os.path.join(ROOT, "tests", "golden", "m4_elsewhere.npz"))
    table = syn.nsx_like_table()
    S = int(d["e_stride"])

    def args(src, prefix, atmosphere):
        a = list(_integrate_args(src, prefix, atmosphere))
        a[19] = np.ascontiguousarray(a[19][::S])
        return a
    cases = []
    cases.append(("c1", args(c1, "int0_", ()), {}, d["c1"]))
    a = args(c1, "int0_", ()); a[10] = d["c1_beam_params"]; a[26] = 1
    cases.append(("c1 beam 1", a, {}, d["c1_beam"]))
    a = args(m2, "t0_int1_", table); a[10] = d["m2_beam_params"]; a[26] = 3
    cases.append(("m2 beam 3", a, {}, d["m2_beam3"]))
    cases.append(("m2 uniform", args(m2, "t0_int1_", table), {}, d["m2"]))
    a = args(m2, "t0_int1_", table); a[10] = d["m2_var_params"]
    cases.append(("m2 per-cell parameters", a, {}, d["m2_var"]))
    a = args(m4, "int0_", table); a[12] = d["m4_corr_params"]; a[23] = table; a[25] = 2
    cases.append(("m4 per-cell Num4D correction", a, {}, d["m4_corr"]))
    return cases, d


def test_oracle_general_integrator(c1, m2):
    cases, d = _general_cases(c1, m2)
    for name, a, kw, ref in cases:
        status, flux = orc.integrate_general(*a, **kw)
        assert status == 0, name
        assert _pulse_err(flux, ref) < 1e-12, name
    a = cases[0][1]
    status, flux = orc.integrate_general(*a, phase_interpolant="Steffen")
    assert status == 0 and _pulse_err(flux, d["c1_steffen"]) < 1e-12


def _intensity_cases():
    from xpsi_b200 import synthetic as syn
    d = np.load(os.path.join(ROOT, "tests", "golden", "intensity.npz"))
    table = syn.nsx_like_table()
    cases = []
    for atm, name in ((None, "BB"), (table, "Num4D")):
        for opt in (0, 1, 2, 3):
            cases.append(("hot %s beam %d" % (name, opt),
                          (d["energies"], d["mu"], d["local_variables"], atm, 0, 'hot', name, opt),
                          d["hot_%s_beam%d" % (name, opt)]))
        cases.append(("elsewhere %s" % name,
                      (d["energies"], d["mu"], np.ascontiguousarray(d["local_variables"][:, :2]), atm, 0,
                       'elsewhere', name, 0), d["elsewhere_%s" % name]))
    return cases


def _point_err(out, ref):
    """point-wise relative error; points more than 30 decades below the largest carry no weight"""
    return float(np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1e-30 * np.max(np.abs(ref)))))


def test_oracle_intensity_seam():
    for name, a, ref in _intensity_cases():
        assert _point_err(orc.intensity(*a), ref) < 1e-9, name


def _odd_leaf_cases(c1):
    """odd leaf counts (the middle leaf is its own mirror) and N_P != N_L; reference outputs in options.npz"""
    d = np.load(os.path.join(ROOT, "tests", "golden", "options.npz"))
    cases = []
    for nl, nph in ((65, 50), (33, 128)):
        a = list(_integrate_args(c1, "int0_", ()))
        a[19] = np.ascontiguousarray(c1["int0_energies"][::8])
        a[20] = np.linspace(0.0, 2.0 * np.pi, nl)
        a[21] = 2.0 * np.pi * np.linspace(0.0, 1.0, nph)
        cases.append((nl, nph, a, d["odd_%d_%d_azinv" % (nl, nph)], d["odd_%d_%d_general" % (nl, nph)]))
    return cases


def test_oracle_odd_leaf_counts(c1):
    for nl, nph, a, ref_az, ref_gen in _odd_leaf_cases(c1):
        status, flux = orc.integrate(*a)
        assert status == 0 and _pulse_err(flux, ref_az) < 1e-12, (nl, nph)
        status, flux = orc.integrate_general(*a)
        assert status == 0 and _pulse_err(flux, ref_gen) < 1e-12, (nl, nph)


def test_oracle_cubic_interpolant(c1):
    """'Cubic' phase interpolant (cspline_periodic) through both integrators, the energy integrator and the
    marginal likelihood, against the reference with xpsi.set_phase_interpolant('Cubic')."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "options.npz"))
    a = list(_integrate_args(c1, "int0_", ()))
    a[19] = np.ascontiguousarray(c1["int0_energies"][::8])
    s, f = orc.integrate(*a, phase_interpolant="Cubic")
    assert s == 0 and _pulse_err(f, d["cubic_azinv"]) < 1e-11
    s, f = orc.integrate_general(*a, phase_interpolant="Cubic")
    assert s == 0 and _pulse_err(f, d["cubic_general"]) < 1e-11
