#!/usr/bin/env python
"""Golden vectors for the point-wise intensity seam, recorded from the reference's
xpsi.surface_radiation_field.intensity (core.pyx:125-308): blackbody and the Num4D table, hot region with
beaming options 0-3 and the elsewhere extension, on seeded random points inside the table."""
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
from xpsi.surface_radiation_field import intensity  # noqa: E402

rng = np.random.default_rng(77)
n = 400
table = syn.nsx_like_table()
out = {}
out["energies"] = np.ascontiguousarray(10.0 ** rng.uniform(-1.0, 1.0, n))
out["mu"] = np.ascontiguousarray(rng.uniform(0.02, 1.0, n))
lv = np.zeros((n, 7))
lv[:, 0] = rng.uniform(5.6, 6.6, n)
lv[:, 1] = rng.uniform(13.9, 14.8, n)
lv[:, 2:6] = rng.uniform([-0.3, -0.2, -0.5, -0.5], [0.3, 0.2, 0.5, 0.5], (n, 4))
lv[:, 6] = 24.0
out["local_variables"] = np.ascontiguousarray(lv)
for atm, name in ((None, "BB"), (table, "Num4D")):
    for opt in (0, 1, 2, 3):
        out["hot_%s_beam%d" % (name, opt)] = intensity(out["energies"], out["mu"], out["local_variables"], atm,
                                                       0, 'hot', name, opt, 1)
    out["elsewhere_%s" % name] = intensity(out["energies"], out["mu"], np.ascontiguousarray(lv[:, :2]), atm,
                                           0, 'elsewhere', name, 0, 1)
np.savez_compressed(os.path.join(HERE, "intensity.npz"), **out)
print("intensity.npz", os.path.getsize(os.path.join(HERE, "intensity.npz")) // 1024, "KiB")
for k in sorted(out):
    if k.startswith(("hot", "else")):
        print(k, float(out[k].min()), float(out[k].max()))
