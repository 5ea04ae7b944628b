"""Several signals per likelihood and the Everywhere star inside the batched pipeline (VERDICT r01 "next" item 7),
against fixtures recorded from the reference build by tests/golden/make_golden_multi.py:

* two ``xpsi.Signal`` objects (two instruments, two data sets, one shared interstellar column, the second with its
  own instrument phase shift) registered from one photosphere signal -- the loop of xpsi/Likelihood.py:346-420 and
  the sum of :494-500;
* ``Everywhere(time_invariant=True)`` as the star (xpsi/Photosphere.py:531-540, xpsi/Everywhere.py:577-601): one
  phase column, phase-averaged data.

Both run from the parameter vector (embed on the GPU) through the C ABI.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def xti_response(syn):
    m, edges, channels, ch_edges = syn.nicer_like_response(n_in=700, chan_lo=40, chan_hi=150)
    return np.ascontiguousarray(0.35 * m), edges


def test_two_signals_joint_likelihood_from_theta():
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.pipeline import BatchedLikelihood
    g = np.load(os.path.join(GOLDEN, "multi_signal.npz"))
    m2 = np.load(os.path.join(GOLDEN, "m2_stu_nsx.npz"))
    names = [str(n) for n in g["names"]]
    P = g["theta"]
    B = P.shape[0]
    mN, eN = syn.nicer_like_response()[:2]
    mX, eX = xti_response(syn)
    pipe = BatchedLikelihood(member_component=[0, 1], max_rings=64, max_azi=64, n_rays=200, energies=g["energies"],
                             leaves=m2["t0_int0_leaves"], phases=m2["t0_int0_phases"], hot_atm_ext=2,
                             hot_atmosphere=syn.nsx_like_table(), image_order_limit=3, response=mN, energy_edges=eN,
                             counts=g["counts_N"], data_phases=np.linspace(0.0, 1.0, 33),
                             exposure_time=syn.M2_EXPOSURE, max_batch=4)
    pipe.set_extras(attenuation=g["attenuation_N_unit"])
    idx = pipe.add_signal(response=mX, energy_edges=eX, counts=g["counts_X"],
                          data_phases=np.linspace(0.0, 1.0, int(g["xti_bins"]) + 1),
                          exposure_time=float(g["xti_exposure"]), attenuation=g["attenuation_X_unit"])
    assert idx == 1 and len(pipe.signals) == 2
    col = {n: i for i, n in enumerate(names)}
    nh = P[:, [i for n, i in col.items() if n.endswith("column_density")][0]]
    spots = syn.m2_spot_batch(pipe, P[:, :11])
    shifts = np.zeros((B, 2))
    shifts[:, 1] = P[:, col["X__phase_shift"]]
    pipe.upload_extras(B, att_power=nh)
    pipe.upload_signal_shifts(B, shifts)
    lnL, status = pipe.eval_spots(spots)
    assert (status == 0).all(), status
    fN, eN_, lN = pipe.fetch_signal(0, B)
    fX, eX_, lX = pipe.fetch_signal(1, B)
    print()
    for b in range(B):
        print("theta %d  joint %.8f (ref %.8f, diff %.2e)  N diff %.2e  X diff %.2e  expected N %.1e X %.1e"
              % (b, lnL[b], g["lnL_total"][b], lnL[b] - g["lnL_total"][b], lN[b] - g["lnL_N"][b],
                 lX[b] - g["lnL_X"][b], rel_err(eN_[b], g["expected_N"][b]), rel_err(eX_[b], g["expected_X"][b])))
    assert np.max(np.abs(lnL - g["lnL_total"])) < 1e-6
    assert np.max(np.abs(lN - g["lnL_N"])) < 1e-6
    assert np.max(np.abs(lX - g["lnL_X"])) < 1e-6
    assert np.allclose(lnL, lN + lX, rtol=0, atol=1e-9)
    for b in range(B):
        assert rel_err(eN_[b], g["expected_N"][b]) < 1e-8
        assert rel_err(eX_[b], g["expected_X"][b]) < 1e-8
    # the sweep path carries the per-signal shifts and the column density too
    lnL_s, st_s = pipe.sweep_spots(spots, att_power=nh, signal_shifts=shifts)
    assert (st_s == 0).all()
    assert np.max(np.abs(lnL_s - g["lnL_total"])) < 1e-6
    # without the instrument phase shift the second signal's term changes, the first one's does not
    pipe.upload_signal_shifts(B, None)
    pipe.upload_extras(B, att_power=nh)
    lnL0, _ = pipe.eval_spots(spots)
    _, _, lN0 = pipe.fetch_signal(0, B)
    assert np.max(np.abs(lN0 - g["lnL_N"])) < 1e-6
    assert np.min(np.abs(lnL0 - g["lnL_total"])) > 1e-3


def test_everywhere_star_in_the_pipeline_from_theta():
    from xpsi_b200 import synthetic as syn
    from xpsi_b200.pipeline import BatchedLikelihood
    g = np.load(os.path.join(GOLDEN, "everywhere.npz"))
    names = [str(n) for n in g["names"]]
    P = g["theta"]
    B = P.shape[0]
    m, e = syn.nicer_like_response()[:2]
    pipe = BatchedLikelihood(energies=g["energies"], response=m, energy_edges=e, counts=g["counts"],
                             data_phases=np.array([0.0, 1.0]), exposure_time=float(g["exposure"]), max_batch=4)
    assert pipe.time_invariant
    pipe.set_extras(everywhere=dict(sqrt_num_cells=24, num_rays=300, atm_ext=2, atmosphere=syn.nsx_like_table(),
                                    image_order_limit=3), attenuation=g["attenuation_unit"])
    col = {n: i for i, n in enumerate(names)}
    spots = pipe.new_spot_batch(B, 300.0)
    spots.set_spacetime(P[:, col["mass"]], P[:, col["radius"]], P[:, col["distance"]], P[:, col["cos_inclination"]], 300.0)
    pipe.upload_extras(B, else_temperature=P[:, col["temperature"]], att_power=P[:, col["column_density"]])
    lnL, status = pipe.eval_spots(spots)
    assert (status == 0).all(), status
    spec = pipe.fetch_elsewhere(B)
    folded, expected, l0 = pipe.fetch_signal(0, B)
    print()
    for b in range(B):
        print("theta %d  lnL %.8f (ref %.8f, diff %.2e)  spectrum %.1e  folded %.1e  expected %.1e"
              % (b, lnL[b], g["lnL_total"][b], lnL[b] - g["lnL_total"][b], rel_err(spec[b], g["spectrum"][b]),
                 rel_err(folded[b, 0, :, 0], g["folded"][b]), rel_err(expected[b], g["expected"][b])))
    for b in range(B):
        assert rel_err(spec[b], g["spectrum"][b]) < 1e-8
        assert rel_err(folded[b, 0, :, 0], g["folded"][b]) < 1e-8
        assert rel_err(expected[b], g["expected"][b]) < 1e-8
    assert np.max(np.abs(lnL - g["lnL_total"])) < 1e-6
    lnL_s, st_s = pipe.sweep_spots(spots, att_power=P[:, col["column_density"]],
                                   else_temperature=P[:, col["temperature"]])
    assert (st_s == 0).all() and np.max(np.abs(lnL_s - g["lnL_total"])) < 1e-6
