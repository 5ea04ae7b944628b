#!/usr/bin/env python
"""Golden vectors for the integrator's optional branches, recorded by calling the reference's
integrator_for_azimuthal_invariance.integrate directly on fixture inputs:
  * disc occultation (R_in < 1e6, common_functions.pyx:110-138) on the C1 blackbody inputs
  * beaming options 1, 2 and 3 (hot_wrapper.pyx:155-199; 3 with nimu = 24) on C1 (BB) and on an M2 member (Num4D)
  * the Steffen phase interpolant (tools/core.pyx:34-52)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_env  # noqa: E402
from xpsi_b200 import synthetic as syn  # noqa: E402

xpsi = ref_env.import_reference()
from xpsi.cellmesh.integrator_for_azimuthal_invariance import integrate  # noqa: E402


def args_of(d, p, atmosphere, params=None, beam_opt=0):
    g = lambda k: d[p + k]
    return [1, float(g("R")), float(g("omega")), float(g("r_s")), float(g("inclination")), g("cellArea"),
            g("radialCoords_of_parallels"), g("r_s_over_r"), g("theta"), g("phi"),
            g("srcCellParams") if params is None else params, g("CELL_RADIATES"), None, int(g("numRays")),
            g("deflection"), g("cos_alpha"), g("lag"), g("maxDeflection"), g("cos_gammaArray"), g("energies"),
            g("leaves"), g("phases"), atmosphere, (), int(g("hot_atm_ext")), 1, beam_opt,
            int(g("image_order_limit"))]


def beam_params(base):
    out = np.zeros(base.shape[:2] + (7,))
    out[..., :2] = base
    out[..., 2:6] = [0.15, -0.08, 0.3, 0.5]
    out[..., 6] = 24.0
    return np.ascontiguousarray(out)


c1 = np.load(os.path.join(HERE, "c1_st_bb.npz"))
m2 = np.load(os.path.join(HERE, "m2_stu_nsx.npz"))
table = syn.nsx_like_table()
out = {}
# the C1 spot sits in the northern hemisphere (never blocked); flip it south so the disc matters
a = args_of(c1, "int0_", ())
th = np.pi - c1["int0_theta"]
a[8] = np.ascontiguousarray(th)
out["disk_theta"] = a[8]
for R_in in (2.0e4, 5.0e4):
    s, f = integrate(*a, R_in=R_in)
    assert s == 0
    out["disk_flux_%d" % int(R_in)] = np.array(f)
s, f = integrate(*a)
out["disk_flux_none"] = np.array(f)
out["beam_params_c1"] = beam_params(c1["int0_srcCellParams"])
out["beam_params_m2"] = beam_params(m2["t0_int1_srcCellParams"])
for opt in (1, 2, 3):
    s, f = integrate(*args_of(c1, "int0_", (), out["beam_params_c1"], opt))
    assert s == 0
    out["beam%d_c1" % opt] = np.array(f)
    s, f = integrate(*args_of(m2, "t0_int1_", table, out["beam_params_m2"], opt))
    assert s == 0
    out["beam%d_m2" % opt] = np.array(f)
xpsi.set_phase_interpolant('Steffen')
s, f = integrate(*args_of(c1, "int0_", ()))
out["steffen_c1"] = np.array(f)
xpsi.set_phase_interpolant('Akima')
np.savez_compressed(os.path.join(HERE, "options.npz"), **out)
print("options.npz", os.path.getsize(os.path.join(HERE, "options.npz")) // 1024, "KiB",
      "disc effect", float(np.max(np.abs(out["disk_flux_20000"] / out["disk_flux_none"].max() - out["disk_flux_none"] / out["disk_flux_none"].max()))))

# ---- odd number of leaves (the middle leaf is its own mirror, pyx:416-419) and N_P != N_L, both integrators ----
from xpsi.cellmesh.integrator import integrate as integrate_general  # noqa: E402
out = dict(np.load(os.path.join(HERE, "options.npz")))
for nl, nph in ((65, 50), (33, 128)):
    a = args_of(c1, "int0_", ())
    a[19] = np.ascontiguousarray(c1["int0_energies"][::8])            # 16 energies keep the fixture small
    a[20] = np.linspace(0.0, 2.0 * np.pi, nl)
    a[21] = 2.0 * np.pi * np.linspace(0.0, 1.0, nph)
    s, f = integrate(*a)
    assert s == 0
    out["odd_%d_%d_azinv" % (nl, nph)] = np.array(f)
    s, f = integrate_general(*a)
    assert s == 0
    out["odd_%d_%d_general" % (nl, nph)] = np.array(f)
np.savez_compressed(os.path.join(HERE, "options.npz"), **out)
print("options.npz (+odd leaf counts)", os.path.getsize(os.path.join(HERE, "options.npz")) // 1024, "KiB")

# ---- 'Cubic' interpolants (tools/core.pyx:84-116: cspline_periodic in phase, cspline in energy) ----------
from xpsi.tools import energy_integrator, energy_interpolator, phase_integrator, phase_interpolator  # noqa: E402
from xpsi.likelihoods.default_background_marginalisation import eval_marginal_likelihood  # noqa: E402
out = dict(np.load(os.path.join(HERE, "options.npz")))
xpsi.set_phase_interpolant('Cubic')
xpsi.set_energy_interpolant('Cubic')
try:
    a = args_of(c1, "int0_", ())
    a[19] = np.ascontiguousarray(c1["int0_energies"][::8])
    s, f = integrate(*a); assert s == 0
    out["cubic_azinv"] = np.array(f)
    s, f = integrate_general(*a); assert s == 0
    out["cubic_general"] = np.array(f)
    sig = np.ascontiguousarray(c1["int0_flux"] / c1["d_sq"])
    out["cubic_eint"] = energy_integrator(1, sig, c1["eint_log10_energies"], c1["eint_log10_edges"])
    comp = c1["marg_components_0"]
    pulse = np.ascontiguousarray(comp[::8])
    out["cubic_pint"] = phase_integrator(1000.0, c1["marg_phases"], pulse, c1["marg_component_phases_0"], 0.37)
    out["cubic_pitp"] = phase_interpolator(np.linspace(0.0, 1.0, 41), c1["marg_component_phases_0"], pulse, -0.2)
    flux4 = np.ascontiguousarray(c1["int0_flux"][:, ::4])
    log10E = np.log10(c1["int0_energies"])
    out["cubic_new_E"] = np.linspace(log10E[0], log10E[-1], 57)
    out["cubic_eitp"] = energy_interpolator(1, flux4, log10E, out["cubic_new_E"])
    res = eval_marginal_likelihood(float(c1["marg_exposure_time"]), c1["marg_phases"], c1["marg_counts"], (comp,),
                                   (c1["marg_component_phases_0"],), c1["marg_phase_shifts"], c1["marg_precomp"],
                                   c1["marg_support"], 1000, 0.0, 1.0e-8, 1.0e-3, 10.0, -1.0e90)
    out["cubic_lnL"] = np.asarray(res[0])
    print("Cubic: lnL", res[0])
finally:
    xpsi.set_phase_interpolant('Akima')
    xpsi.set_energy_interpolant('Steffen')
np.savez_compressed(os.path.join(HERE, "options.npz"), **out)
print("options.npz (+Cubic interpolants)", os.path.getsize(os.path.join(HERE, "options.npz")) // 1024, "KiB")
