#!/usr/bin/env python
"""Golden vectors for the general (no azimuthal invariance) integrator, recorded by calling the
reference's xpsi.cellmesh.integrator.integrate (integrator.pyx:48-667) directly on fixture inputs:
  * c1:      the C1 blackbody member (what HotRegion(symmetry=False) passes)
  * c1_beam: the same with beaming option 1 (hot_wrapper.pyx:155-172)
  * c1_steffen: the same with the Steffen phase interpolant for the GEOM spline
  * m2:      an M2 member with the Num4D table (parameters uniform over the mesh)
  * m2_var:  the same member with log T and log g varying from cell to cell
  * m4_corr: an M4 member with the Num4D elsewhere correction, correction parameters varying per cell
Energies are thinned to 32 to keep the CPU run short; everything else is as in the fixtures."""
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
from xpsi.cellmesh.integrator import integrate  # noqa: E402

E_STRIDE = 4


def args_of(d, p, atmosphere, params=None, beam_opt=0, correction=None, els_atmosphere=(), else_ext=1):
    g = lambda k: d[p + k]
    return [1, float(g("R")), float(g("omega")), float(g("r_s")), float(g("inclination")), g("cellArea"),
            g("radialCoords_of_parallels"), g("r_s_over_r"), g("theta"), g("phi"),
            g("srcCellParams") if params is None else params, g("CELL_RADIATES"), correction, int(g("numRays")),
            g("deflection"), g("cos_alpha"), g("lag"), g("maxDeflection"), g("cos_gammaArray"),
            np.ascontiguousarray(g("energies")[::E_STRIDE]),
            g("leaves"), g("phases"), atmosphere, els_atmosphere, int(g("hot_atm_ext")), else_ext, beam_opt,
            int(g("image_order_limit"))]


def run(a):
    s, f = integrate(*a)
    assert s == 0
    return np.array(f)


c1 = np.load(os.path.join(HERE, "c1_st_bb.npz"))
m2 = np.load(os.path.join(HERE, "m2_stu_nsx.npz"))
m4 = np.load(os.path.join(HERE, "m4_elsewhere.npz"))
table = syn.nsx_like_table()
out = {"e_stride": np.int64(E_STRIDE)}

out["c1"] = run(args_of(c1, "int0_", ()))
bp = np.zeros(c1["int0_srcCellParams"].shape[:2] + (7,))
bp[..., :2] = c1["int0_srcCellParams"]
bp[..., 2:6] = [0.15, -0.08, 0.3, 0.5]
bp[..., 6] = 24.0
out["c1_beam_params"] = np.ascontiguousarray(bp)
out["c1_beam"] = run(args_of(c1, "int0_", (), out["c1_beam_params"], 1))
bm = np.zeros(m2["t0_int1_srcCellParams"].shape[:2] + (7,))
bm[..., :2] = m2["t0_int1_srcCellParams"]
bm[..., 2:6] = [0.15, -0.08, 0.3, 0.5]
bm[..., 6] = 24.0
out["m2_beam_params"] = np.ascontiguousarray(bm)
out["m2_beam3"] = run(args_of(m2, "t0_int1_", table, out["m2_beam_params"], 3))
xpsi.set_phase_interpolant('Steffen')
out["c1_steffen"] = run(args_of(c1, "int0_", ()))
xpsi.set_phase_interpolant('Akima')

out["m2"] = run(args_of(m2, "t0_int1_", table))
par = np.array(m2["t0_int1_srcCellParams"])
phi = m2["t0_int1_phi"]
par[..., 0] += 0.04 * np.sin(3.0 * phi) - 0.02          # log10 T over the spot
par[..., 1] += 0.03 * np.cos(2.0 * phi)                  # log10 g
out["m2_var_params"] = np.ascontiguousarray(par)
out["m2_var"] = run(args_of(m2, "t0_int1_", table, out["m2_var_params"]))

corr = np.array(m4["int0_correction_srcCellParams"])
corr[..., 0] += 0.03 * np.cos(m4["int0_phi"])
out["m4_corr_params"] = np.ascontiguousarray(corr)
out["m4_corr"] = run(args_of(m4, "int0_", table, correction=out["m4_corr_params"], els_atmosphere=table,
                             else_ext=2))
np.savez_compressed(os.path.join(HERE, "general.npz"), **out)
print("general.npz", os.path.getsize(os.path.join(HERE, "general.npz")) // 1024, "KiB")
for k in ("c1", "c1_beam", "c1_steffen", "m2", "m2_var", "m4_corr"):
    print(k, out[k].shape, float(out[k].max()))
