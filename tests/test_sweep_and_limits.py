"""GPU tests of (i) the device-resident sweep behind ``xpsi_b200.sampling`` (config 5 of BASELINE.json: many
distinct parameter vectors, blocks of ``max_batch``) and (ii) every capability limit the kernels enforce: a
configuration outside the coverage must be refused loudly (exception or status 3), never answered wrongly."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _pipeline(max_batch, **over):
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.pipeline import BatchedLikelihood
    m2 = np.load(os.path.join(GOLDEN, "m2_stu_nsx.npz"))
    matrix, edges = syn.nicer_like_response()[:2]
    kw = dict(member_component=[0, 1], max_rings=64, max_azi=64, n_rays=200, energies=m2["t0_int0_energies"],
              leaves=m2["t0_int0_leaves"], phases=m2["t0_int0_phases"], hot_atm_ext=2,
              hot_atmosphere=syn.nsx_like_table(), image_order_limit=3, response=matrix, energy_edges=edges,
              counts=m2["counts"], data_phases=np.linspace(0.0, 1.0, 33), exposure_time=syn.M2_EXPOSURE,
              max_batch=max_batch)
    kw.update(over)
    return BatchedLikelihood(**kw)


def test_sweep_equals_blockwise_evaluation_and_fixture():
    """700 distinct parameter vectors: one upload + blocks of 256 + one download gives what 700 rows evaluated
    block by block through ``eval_spots`` give, in row order, including a ragged last block; the rows are the
    head of the reference fixture, so status and lnL are also checked against the reference."""
    from xpsi_b200 import sampling
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.likelihood import Likelihood
    pipe = _pipeline(256)
    P = syn.m2_bench_thetas(0, 700)
    lnL_s, st_s = pipe.sweep_spots(syn.m2_spot_batch(pipe, P))
    lnL_b = np.empty(700)
    st_b = np.empty(700, dtype=np.int32)
    for i in range(0, 700, 256):
        blk = P[i:i + 256]
        lnL_b[i:i + len(blk)], st_b[i:i + len(blk)] = pipe.eval_spots(syn.m2_spot_batch(pipe, blk))
    assert np.array_equal(st_s, st_b)
    ok = st_s == 0
    assert ok.sum() > 300
    d = np.abs(lnL_s[ok] - lnL_b[ok]) / np.abs(lnL_b[ok])
    print("sweep vs blockwise: max rel diff %.2e" % d.max())
    assert d.max() < 1.0e-13                                   # same kernels; ring sums use fp64 atomics
    fx = np.load(os.path.join(GOLDEN, "m2_prior.npz"))
    assert ((st_s[:512] != 0) == fx["early_exit"][:512]).all()
    okf = ~fx["early_exit"][:512]
    rel = np.abs(lnL_s[:512][okf] - fx["lnL"][:512][okf]) / np.abs(fx["lnL"][:512][okf])
    assert rel.max() < 3.0e-8
    # partial ranges and the Likelihood / sampling front ends
    pipe.sweep_run(300, 50)
    l2, s2 = pipe.sweep_download(300, 50)
    assert np.array_equal(s2, st_s[300:350]) and np.allclose(l2[s2 == 0], lnL_s[300:350][s2 == 0], rtol=1e-13)
    like = Likelihood(pipe, lambda p, X: syn.m2_spot_batch(p, X))
    l3, s3 = like.sweep_local(P[:300])
    assert np.array_equal(s3, st_s[:300]) and np.all(np.isnan(l3[s3 != 0]))
    assert np.allclose(l3[s3 == 0], lnL_s[:300][s3 == 0], rtol=1e-13)
    info = {}
    l4, s4 = sampling.sweep(like, P, info=info)
    assert info["rows"] == 700 and info["device_ms"] > 0.0
    assert np.array_equal(s4, st_s) and np.allclose(l4[ok], lnL_s[ok], rtol=1e-13)
    loglike = sampling.vectorized_loglike(like, syn.m2_prior())
    out = loglike(P[:64])
    good = st_s[:64] == 0
    assert np.allclose(out[good], lnL_s[:64][good], rtol=1e-13)
    assert np.all(out[~good] <= -1.0e89)                       # early exits: random value near llzero


def test_capability_limits_are_refused_loudly():
    from xpsi_b200 import _lib
    from xpsi_b200 import synthetic as syn
    from xpsi_b200 import tools
    from xpsi_b200.likelihood import Likelihood
    refused = (_lib.XpsiB200Error, NotImplementedError)
    m2 = np.load(os.path.join(GOLDEN, "m2_stu_nsx.npz"))
    P = syn.m2_bench_thetas(0, 4)
    # more than 128 data phase bins (k_marginal) / more than 128 output phases (k_azinv_flux)
    with pytest.raises(refused):
        _pipeline(4, data_phases=np.linspace(0.0, 1.0, 130), counts=np.ones((270, 129)))
    with pytest.raises(refused):
        pipe = _pipeline(4, phases=np.linspace(0.0, 2.0 * np.pi, 129), leaves=np.linspace(0.0, 2.0 * np.pi, 129))
        pipe.eval_spots(syn.m2_spot_batch(pipe, P))
    # beaming needs the five beaming parameters behind (log T, log g): n_params >= 7
    with pytest.raises(refused):
        pipe = _pipeline(4)
        pipe.set_extras(beam_opt=1)
        pipe.eval_spots(syn.m2_spot_batch(pipe, P))
    # a padded mesh smaller than the embed's max_sqrt_num_cells
    with pytest.raises(ValueError):
        syn.m2_spot_batch(_pipeline(4, max_rings=48, max_azi=48), P)
    # Doppler spread beyond the slab budget (kDopplerDex): a 1 kHz, 16 km star is refused with status 3, and the
    # Likelihood front end turns that into an exception instead of a "zero likelihood" point
    pipe = _pipeline(4)
    fast = np.array(P)
    fast[:, 1] = 16.0
    fast[:, 5] = 1.4                       # primary spot next to the equator: the fastest rings of the star
    sb = syn.m2_spot_batch(pipe, fast)
    sb.set_spacetime(fast[:, 0], fast[:, 1], fast[:, 2], fast[:, 3], 1200.0)
    sb.mode_frequency = 1200.0
    lnL, status = pipe.eval_spots(sb)
    print("1.2 kHz / 16 km star: status", status)
    assert (status == _lib.EUNSUPPORTED).all()

    def fill(p, X):
        s = syn.m2_spot_batch(p, X)
        s.set_spacetime(X[:, 0], X[:, 1], X[:, 2], X[:, 3], 1200.0)
        s.mode_frequency = 1200.0
        return s
    like = Likelihood(pipe, fill)
    with pytest.raises(NotImplementedError):
        like(fast[0])
    with pytest.raises(NotImplementedError):
        like.batch(fast)
    lnL, status = like.batch(fast, strict=False)
    assert np.all(np.isnan(lnL)) and (status == 3).all()
    # the global C2 phase spline keeps its tridiagonal system in per-thread memory: at most 260 nodes
    tools.set_phase_interpolant('Cubic')
    try:
        x = np.linspace(0.0, 1.0, 300)
        with pytest.raises(refused):
            tools.phase_interpolator(np.linspace(0.0, 1.0, 16), x, np.ones((2, 300)), 0.0)
    finally:
        tools.set_phase_interpolant('Akima')
    # more image orders than the kernels keep (kMaxImages = 6)
    from xpsi_b200.cellmesh.integrator_for_azimuthal_invariance import integrate
    c1 = np.load(os.path.join(GOLDEN, "c1_st_bb.npz"))
    g = lambda k: c1["int0_" + k]
    args = [1, float(g("R")), float(g("omega")), float(g("r_s")), float(g("inclination")), g("cellArea"),
            g("radialCoords_of_parallels"), g("r_s_over_r"), g("theta"), g("phi"), g("srcCellParams"),
            g("CELL_RADIATES"), None, int(g("numRays")), g("deflection"), g("cos_alpha"), g("lag"),
            g("maxDeflection"), g("cos_gammaArray"), g("energies"), g("leaves"), g("phases"), (), (), 1, 1, 0, 7]
    with pytest.raises(refused):
        integrate(*args)
    # per-batch extras belong to the batch size they were uploaded for: evaluating another size is refused
    # instead of reading stale attenuation powers
    pipe = _pipeline(4)
    pipe.set_extras(attenuation=np.exp(-0.3 * (0.2025 + 0.005 * np.arange(1500)) ** -2.5))
    pipe.upload_extras(2, att_power=np.array([0.5, 0.7]))
    pipe.eval_spots(syn.m2_spot_batch(pipe, P[:2]))
    with pytest.raises(refused):
        pipe.eval_spots(syn.m2_spot_batch(pipe, P))


def test_deterministic_mode_is_bitwise_reproducible():
    """Two-stage ordered ring reduction: with it lnL is bitwise equal from run to run and does
    not depend on where a parameter vector sits in its batch; it agrees with the default (fp64 atomics) to
    rounding.  Rings wider than the tile budget (polar caps) go through the scalar flux kernel in both modes."""
    from xpsi_b200 import synthetic as syn
    pipe = _pipeline(256)
    P = syn.m2_bench_thetas(0, 256)
    spots = lambda X: syn.m2_spot_batch(pipe, X)
    lnL_a, st_a = pipe.eval_spots(spots(P))
    pipe.set_deterministic(True)
    lnL_1, st_1 = pipe.eval_spots(spots(P))
    flux_1 = pipe.fetch(256, folded=False, expected=False)[0]
    lnL_2, st_2 = pipe.eval_spots(spots(P))
    flux_2 = pipe.fetch(256, folded=False, expected=False)[0]
    assert np.array_equal(st_1, st_a) and np.array_equal(st_2, st_a)
    ok = st_a == 0
    assert np.array_equal(flux_1, flux_2), "member signals differ between two deterministic runs"
    assert np.array_equal(lnL_1[ok], lnL_2[ok])
    perm = np.random.default_rng(1).permutation(256)
    lnL_p, st_p = pipe.eval_spots(spots(P[perm]))
    assert np.array_equal(st_p, st_a[perm])
    assert np.array_equal(lnL_p[ok[perm]], lnL_1[perm][ok[perm]]), "lnL depends on the position in the batch"
    rel = np.abs(lnL_1[ok] - lnL_a[ok]) / np.abs(lnL_a[ok])
    print("deterministic vs atomic ring sums: max rel diff %.2e" % rel.max())
    assert rel.max() < 1.0e-13
    pipe.set_deterministic(False)
    lnL_b, st_b = pipe.eval_spots(spots(P))
    assert np.allclose(lnL_b[ok], lnL_a[ok], rtol=1e-13)


def test_chunks_handed_back_to_the_scalar_kernel_give_the_same_signal():
    """The tensor-core flux kernel hands a (ring, energy chunk) back to the scalar kernel when GSL's Akima takes
    different slopes on the two sides of a node (two exactly straight segments meeting: never for a physical
    profile).  XPSI_B200_FORCE_REDO=1 hands *every* chunk back: the member signals must agree with the tensor-core
    result to rounding -- no chunk lost, none counted twice."""
    from xpsi_b200 import synthetic as syn
    P = syn.m2_bench_thetas(0, 6)
    pipe = _pipeline(8)
    spots = syn.m2_spot_batch(pipe, P)
    lnL_a, st_a = pipe.eval_spots(spots)
    flux_a = pipe.fetch(6, folded=False, expected=False)[0]
    os.environ["XPSI_B200_FORCE_REDO"] = "1"
    try:
        lnL_b, st_b = pipe.eval_spots(spots)
        flux_b = pipe.fetch(6, folded=False, expected=False)[0]
    finally:
        del os.environ["XPSI_B200_FORCE_REDO"]
    assert np.array_equal(st_a, st_b)
    rel = np.abs(flux_a - flux_b).max() / np.abs(flux_a).max()
    print("tensor-core vs handed-back (scalar kernel) member signals: max rel diff %.2e; lnL diff %.2e"
          % (rel, np.nanmax(np.abs(lnL_a - lnL_b))))
    assert rel < 1e-12
    ok = st_a == 0
    assert np.max(np.abs(lnL_a[ok] - lnL_b[ok])) < 1e-7
