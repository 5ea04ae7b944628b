"""``xpsi_b200.from_xpsi``: the GPU likelihood built from the reference's own constructed objects (row a15 of
SURVEY.md s8: ``xpsi.Likelihood(star, signals)`` compatibility).  The reference build in ``oracle/_ref`` provides
the objects (stock constructors, the model set-ups of tests/golden/make_golden*.py) and, on the GPU box, the
checker: ``from_xpsi(like)(p)`` against the lnL the unmodified reference recorded for the same vector.

CPU part: the static read-out and the per-row parameter walk (derived parameters, ceding members, per-region
cell budgets) against the reference objects' own values.
"""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, GOLDEN)
    import ref_env
    if not ref_env.available():
        pytest.skip("oracle/_ref (reference build) is not on this box")
    with contextlib.redirect_stdout(io.StringIO()):
        ref_env.import_reference()
        import make_golden as mg
        import make_golden_m3 as m3
        import make_golden_m4 as m4
        import make_golden_multi as mm
    return mg, m3, m4, mm


class FakePipe:
    """host-only stand-in for BatchedLikelihood.new_spot_batch (the walk does not touch the device)"""
    time_invariant = False

    def __init__(self, M, C):
        self.n_members, self.n_components = M, C
        self.shape = dict(max_rings=128, max_azi=128)

    def new_spot_batch(self, B, mode_frequency, **kw):
        from xpsi_b200.pipeline import SpotBatch
        return SpotBatch(B, self.n_members, self.n_components, mode_frequency, **kw)


def test_parameter_walk_matches_the_reference_objects(ref):
    mg, m3, m4, mm = ref
    from xpsi_b200 import from_xpsi as fx
    from xpsi_b200 import synthetic as syn
    with contextlib.redirect_stdout(io.StringIO()):
        like = m3.build_m3(mg.Recorder(), np.ones((270, 32)))[0]
    model = fx._Model(like, True)
    g = np.load(os.path.join(GOLDEN, "m3_cst_pdt.npz"))
    assert [str(n) for n in g["names"]] == model.names
    P = np.vstack([g["theta"], g["theta"]])
    P[1, model.names.index("mass")] = 1.6
    # CST primary (superseding + omission) and PDT secondary (superseding + ceding): 3 members, 2 components
    assert model.member_region == [0, 1, 1] and model.member_is_cede == [0, 0, 1]
    spots, extras = model.fill(FakePipe(3, 2), P)
    assert extras == {}
    hot = like.star.photospheres[0].hot.objects
    like_set = fx._subspace_call(like)
    for b in range(2):
        like_set(P[b])
        p, s = hot
        assert spots.colatitude[b, 0] == p['super_colatitude'] and spots.ang_radius[b, 0] == p['super_radius']
        assert spots.hole_radius[b, 0] == p['omit_radius'] and spots.hole_colatitude[b, 0] == p['omit_colatitude']
        assert spots.temperature[b, 1] == s['super_temperature'] and spots.temperature[b, 2] == s['cede_temperature']
        assert spots.ang_radius[b, 2] == s['cede_radius'] and spots.hole_radius[b, 2] == s['super_radius']
        assert spots.phase_shifts[b, 0] == p['phase_shift'] and spots.phase_shifts[b, 1] == s['phase_shift']
        assert abs(spots.phi_shift[b, 2] - (s['cede_azimuth'] + (np.pi if s._is_antiphased else 0.0))) < 1e-15
        st = like.star.spacetime
        assert spots.R_eq[b] == st.R and spots.r_s[b] == st.r_s and abs(spots.zeta[b] - st.zeta) < 1e-15
        assert abs(spots.epsilon[b] - st.epsilon) < 1e-15 * st.epsilon + 1e-18 and spots.d_sq[b] == st.d_sq
    assert list(spots.partner) == [-1, 2, 1] and list(spots.is_cede) == [0, 0, 1]
    for m, h in zip(range(3), (hot[0], hot[1], hot[1])):
        assert tuple(spots.member_cells[m]) == (h._num_cells, h._min_sqrt_num_cells, h._max_sqrt_num_cells)
    assert spots.mode_frequency == syn.M2_FREQUENCY


def test_interstellar_probe_and_signal_readout(ref):
    mg, m3, m4, mm = ref
    from xpsi_b200 import from_xpsi as fx
    g = np.load(os.path.join(GOLDEN, "multi_signal.npz"))
    with contextlib.redirect_stdout(io.StringIO()):
        like, sigN, sigX = mm.build_two_signals(g["counts_N"], g["counts_X"])
    model = fx._Model(like, True)
    assert len(model.signals) == 2
    model.interstellar = None
    kwN, attN = model._signal_kwargs(sigN)
    kwX, attX = model._signal_kwargs(sigX)
    assert np.allclose(attN, g["attenuation_N_unit"], rtol=1e-14) and np.allclose(attX, g["attenuation_X_unit"], rtol=1e-14)
    assert kwX["response"].shape == (110, 700) and kwX["counts"].shape == (110, 16) and kwX["exposure_time"] == 4.0e5
    spots, extras = model.fill(FakePipe(2, 2), g["theta"])
    col = [str(n) for n in g["names"]]
    assert np.array_equal(extras["att_power"], g["theta"][:, [i for i, n in enumerate(col) if n.endswith("column_density")][0]])
    assert np.array_equal(extras["signal_shifts"][:, 1], g["theta"][:, col.index("X__phase_shift")])
    assert np.all(extras["signal_shifts"][:, 0] == 0.0)
    # the secondary's temperature is a Derive instance: primary - 0.2
    assert np.allclose(spots.temperature[:, 1], spots.temperature[:, 0] - 0.2, rtol=0, atol=1e-15)


# ------------------------------------------------------------------------------------------- GPU: lnL parity
def _check(gpu, P, ref_lnL, tag, tol=1e-6):
    P = np.atleast_2d(P)
    lnL, status = gpu.batch(P)
    assert (status == 0).all(), (tag, status)
    d = np.abs(lnL - np.atleast_1d(ref_lnL))
    print("%-14s from_xpsi(like).batch vs the reference's recorded lnL: max |diff| %.2e  (lnL %s)"
          % (tag, d.max(), np.array2string(lnL, precision=6)))
    assert d.max() < tol
    # the scalar call (xpsi/Likelihood.py:450-511 signature) goes through the same pipeline
    v = gpu(P[0])
    assert abs(v - float(np.atleast_1d(ref_lnL)[0])) < tol


@pytest.mark.gpu
def test_from_xpsi_every_configuration(ref):
    mg, m3, m4, mm = ref
    from xpsi_b200 import from_xpsi
    rec = mg.Recorder()
    quiet = contextlib.redirect_stdout(io.StringIO())
    print()
    # config 1: examples_fast ST blackbody at the published known-answer point
    with quiet:
        like = mg.build_c1(rec)[0]
    c1 = np.load(os.path.join(GOLDEN, "c1_st_bb.npz"))
    gpu = from_xpsi.from_xpsi(like, max_batch=4)
    _check(gpu, c1["theta"], c1["lnL_total"], "C1 ST BB")
    assert abs(gpu(c1["theta"]) + 47881.27817666349) < 1e-5 * 47881.27817666349
    # config 2: ST-U NSX (secondary temperature derived, regions with different max_sqrt_num_cells)
    m2 = np.load(os.path.join(GOLDEN, "m2_stu_nsx.npz"))
    with quiet:
        like = mg.build_m2(rec, m2["counts"])[0]
    gpu = from_xpsi.from_xpsi(like, max_batch=4)
    _check(gpu, np.vstack([m2["t0_theta"], m2["t1_theta"]]), [m2["t0_lnL_total"], m2["t1_lnL_total"]], "M2 ST-U NSX")
    # config 3: CST + PDT (omission hole, superseding + ceding members)
    g = np.load(os.path.join(GOLDEN, "m3_cst_pdt.npz"))
    with quiet:
        like = m3.build_m3(rec, g["counts"])[0]
    _check(from_xpsi.from_xpsi(like, max_batch=4), g["theta"], g["lnL_total"], "M3 CST+PDT")
    # config 4a: Elsewhere + interstellar
    g = np.load(os.path.join(GOLDEN, "m4_elsewhere.npz"))
    with quiet:
        like = m4.build_m4(rec, g["counts"], [])[0]
    _check(from_xpsi.from_xpsi(like, max_batch=4), g["theta"], g["lnL_total"], "M4 Elsewhere")
    # config 4b: Everywhere(time_invariant=True) + interstellar
    g = np.load(os.path.join(GOLDEN, "everywhere.npz"))
    with quiet:
        like = mm.build_everywhere(g["counts"])[0]
    _check(from_xpsi.from_xpsi(like, max_batch=4), g["theta"], g["lnL_total"], "M4 Everywhere")
    # two instruments
    g = np.load(os.path.join(GOLDEN, "multi_signal.npz"))
    with quiet:
        like = mm.build_two_signals(g["counts_N"], g["counts_X"])[0]
    _check(from_xpsi.from_xpsi(like, max_batch=4), g["theta"], g["lnL_total"], "two signals")


@pytest.mark.gpu
def test_from_xpsi_blocks_overlap_fill_and_evaluation(ref):
    """Several blocks through ``from_xpsi(like).batch``: the parameter walk of block k + 1 runs in a worker thread
    while block k is on the GPU; the result must equal the one of the vectorised fill function, row by row."""
    import time
    mg, m3, m4, mm = ref
    from xpsi_b200 import from_xpsi
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.likelihood import Likelihood
    m2 = np.load(os.path.join(GOLDEN, "m2_stu_nsx.npz"))
    with contextlib.redirect_stdout(io.StringIO()):
        like = mg.build_m2(mg.Recorder(), m2["counts"])[0]
    gpu = from_xpsi.from_xpsi(like, max_batch=512)
    P = syn.m2_bench_thetas(0, 2304)                       # 4.5 blocks
    gpu.batch(P[:512])                                     # warm-up
    t0 = time.perf_counter()
    lnL, st = gpu.batch(P)
    t1 = time.perf_counter()
    direct = Likelihood(gpu.pipeline, lambda pl, X: syn.m2_spot_batch(pl, X))
    t2 = time.perf_counter()
    lnL_d, st_d = direct.batch(P)
    t3 = time.perf_counter()
    print("\nfrom_xpsi(like).batch: %.0f evals/s (parameter walk through the reference's objects, overlapped); "
          "vectorised fill: %.0f evals/s" % (P.shape[0] / (t1 - t0), P.shape[0] / (t3 - t2)))
    assert np.array_equal(st, st_d)
    ok = st == 0
    assert np.max(np.abs(lnL[ok] - lnL_d[ok])) < 1e-7
