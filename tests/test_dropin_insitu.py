"""In-situ drop-in proof (VERDICT r01 "next" item 3): the reference's own objects -- ``xpsi.Likelihood`` over
``Star`` / ``Photosphere`` / ``HotRegion`` / ``Signal`` / ``Instrument`` / ``CustomSignal`` from the build in
``oracle/_ref`` -- are constructed unchanged, the compiled seams are rebound by ``xpsi_b200.dropin.install``
exactly as INTEGRATION.md sections 1-3 describe (xpsi/HotRegion.py:551-572, xpsi/Elsewhere.py:5,
xpsi/Signal.py:10-11, xpsi/Instrument.py:192-197, examples_fast/Modules/CustomSignal.py:7-8), and
``xpsi.Likelihood.__call__`` (xpsi/Likelihood.py:450) is evaluated.  The reference is the checker and the
harness here; every rebound callable computes on the GPU through the C ABI.
"""
import os
import sys
import time

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

KNOWN_ANSWER = -47881.27817666349        # xpsi/tests/test_likelihood.py:134
C1_P = [1.4, 10, 1., np.cos(60 * np.pi / 180), 0.0, 70 * np.pi / 180, 0.75, 6.8]


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, GOLDEN)
    import ref_env
    if not ref_env.available():
        pytest.skip("oracle/_ref (reference build) is not on this box")
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        xpsi = ref_env.import_reference()
        import make_golden as mg
    return xpsi, mg


def _timed(like, p, n=3):
    import contextlib
    import io
    best = 1e30
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(n):
            t0 = time.perf_counter()
            v = like(list(p), force=True)
            best = min(best, time.perf_counter() - t0)
    return float(v), best


def test_reference_likelihood_object_with_gpu_seams(ref):
    xpsi, mg = ref
    from xpsi_b200 import _lib, dropin
    from xpsi_b200 import synthetic as syn
    m2 = np.load(os.path.join(GOLDEN, "m2_stu_nsx.npz"))
    rec = mg.Recorder()                                   # only used by build_* to wrap the integrators
    # ---- unpatched runs: the reference on the CPU ---------------------------------------------------
    like_c1 = mg.build_c1(rec)[0]
    like_m2 = mg.build_m2(rec, m2["counts"])[0]
    c1_cpu, t_c1_cpu = _timed(like_c1, C1_P)
    m2_cpu = [_timed(like_m2, m2["t%d_theta" % t]) for t in range(2)]
    assert abs(c1_cpu - KNOWN_ANSWER) < 1e-5 * abs(KNOWN_ANSWER)
    # ---- rebind the seams, build the SAME models again with the stock constructors ------------------
    k0 = _lib.counters()[0]
    h = dropin.install(xpsi, user_modules=[mg.CustomSignal_module])
    try:
        hot_c1 = mg.build_c1(rec)
        like_c1g = hot_c1[0]
        like_m2g = mg.build_m2(rec, m2["counts"])[0]
        # the instances hold the GPU callables (HotRegion.symmetry setter imported them late)
        from xpsi_b200.cellmesh.integrator_for_azimuthal_invariance import integrate as gpu_integrate
        assert sys.modules["xpsi.cellmesh.integrator_for_azimuthal_invariance"].integrate is gpu_integrate
        assert sys.modules["xpsi.Signal"].energy_integrator.__module__.startswith("xpsi_b200")
        assert mg.CustomSignal_module.eval_marginal_likelihood.__module__.startswith("xpsi_b200")
        c1_gpu, t_c1_gpu = _timed(like_c1g, C1_P)
        m2_gpu = [_timed(like_m2g, m2["t%d_theta" % t]) for t in range(2)]
    finally:
        h.uninstall()
    launches = _lib.counters()[0] - k0
    print()
    print("C1  xpsi.Likelihood  CPU %.10f  GPU seams %.10f  diff %.2e  known answer diff %.2e  (%.0f ms vs %.0f ms per call)"
          % (c1_cpu, c1_gpu, c1_gpu - c1_cpu, c1_gpu - KNOWN_ANSWER, 1e3 * t_c1_cpu, 1e3 * t_c1_gpu))
    for t in range(2):
        print("M2 theta %d xpsi.Likelihood  CPU %.8f  GPU seams %.8f  diff %.2e  (%.0f ms vs %.0f ms per call)"
              % (t, m2_cpu[t][0], m2_gpu[t][0], m2_gpu[t][0] - m2_cpu[t][0], 1e3 * m2_cpu[t][1], 1e3 * m2_gpu[t][1]))
    print("GPU kernels launched through the rebound seams:", launches)
    assert launches > 0
    assert abs(c1_gpu - c1_cpu) < 1e-6
    assert abs(c1_gpu - KNOWN_ANSWER) < 1e-5 * abs(KNOWN_ANSWER)
    for t in range(2):
        assert abs(m2_gpu[t][0] - m2_cpu[t][0]) < 1e-6
        assert abs(m2_gpu[t][0] - float(m2["t%d_lnL_total" % t])) < 1e-6
    # after uninstall the reference names are back
    assert not sys.modules["xpsi.Signal"].energy_integrator.__module__.startswith("xpsi_b200")
    assert sys.modules["xpsi.cellmesh.integrator_for_azimuthal_invariance"].integrate is not gpu_integrate
