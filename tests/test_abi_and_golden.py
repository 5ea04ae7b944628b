"""CPU-side checks: the C-ABI library loads and exports every symbol the header
declares (no compute without a GPU), and the golden fixtures are self-consistent
with the reference's published known answer."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from xpsi_b200 import _lib
    header = open(os.path.join(ROOT, "include", "xpsi_b200.h")).read()
    declared = set(re.findall(r"\b(xpsi_b200_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(_lib.EXPORTED)


def test_no_device_is_a_loud_error_not_a_fallback():
    from xpsi_b200 import _lib
    if _lib.lib.xpsi_b200_device_count() > 0:
        return
    from xpsi_b200.tools import energy_integrator
    x = np.linspace(-1, 1, 16)
    try:
        energy_integrator(1, np.ones((16, 4)), x, np.linspace(-1, 1, 9))
    except _lib.XpsiB200Error as e:
        assert "no usable CUDA device" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("expected a loud failure without a GPU")


def test_golden_c1_matches_published_known_answer(c1):
    # xpsi/tests/test_likelihood.py:134 pins -47881.27817666349 at rtol 1e-5
    assert abs(float(c1["lnL_total"]) + 47881.27817666349) < 1e-5 * 47881.27817666349
    assert abs(float(c1["lnL_total"]) - float(c1["marg_lnL"])) < 1e-9
    assert c1["int0_flux"].shape == (128, 64)


def test_synthetic_workload_is_deterministic():
    from xpsi_b200 import synthetic as syn
    t = syn.nsx_like_table()
    assert t[4].size == 35 * 14 * 67 * 166 and np.all(t[4] > 0)
    a = syn.m2_theta_batch(8)
    b = syn.m2_theta_batch(8)
    assert np.array_equal(a, b)
    m, e, c, ce = syn.nicer_like_response()
    assert m.shape == (270, 1500) and (m.sum(axis=0) > 0).all()
