#!/usr/bin/env python
"""Golden vectors for the models of round 2's pipeline widening, recorded from the reference build:

  multi_signal.npz  ST-U NSX hot regions seen by TWO instruments through two ``xpsi.Signal`` objects that share
                    one interstellar column (docs/source/Instrument_synergy.ipynb; the loop of
                    xpsi/Likelihood.py:346-420): a NICER-like signal (270 channels x 1500 inputs, 32 phase bins)
                    and an "XTI-B" signal (110 channels x 700 inputs, 16 phase bins, its own exposure and a free
                    instrument phase shift that its likelihood call honours, xpsi/Signal.py:221-245,581-583).  Stored: the shared energy array the
                    reference builds (Signal.construct_energy_array), both data sets, and per parameter vector the
                    joint lnL, each signal's log-likelihood and expected counts.
  everywhere.npz    ``Everywhere(time_invariant=True)`` star (Num4D atmosphere, 24 x 24 closed mesh, 300 rays)
                    behind one instrument with phase-averaged data (one phase bin) and interstellar attenuation:
                    the spectrum, the registered (folded) signal and lnL from the parameter vector
                    (xpsi/Photosphere.py:531-540, xpsi/Everywhere.py:577-601).

Usage: python oracle/build_ref.py && python tests/golden/make_golden_multi.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from make_golden_m4 import Interstellar  # noqa: E402

xpsi, syn = mg.xpsi, mg.syn

XTI_EXPOSURE = 4.0e5
XTI_BINS = 16
EV_EXPOSURE = 1.0e6


class ShiftedSignal(mg.CustomSignal):
    """A signal that honours its own instrument phase shift: ``Signal.shifts`` = hot-region shifts +
    ``self['phase_shift']`` (xpsi/Signal.py:581-583).  The example ``CustomSignal`` classes pass ``self._shifts``,
    i.e. they ignore the instrument's parameter; a user who binds it writes exactly this ``__call__``."""

    def __call__(self, *args, **kwargs):
        self.loglikelihood, self.expected_counts, self.background_signal, self.background_given_support = \
            mg.CustomSignal_module.eval_marginal_likelihood(
                self._data.exposure_time, self._data.phases, self._data.counts, self._signals, self._phases,
                self.shifts, self._precomp, self._support, self._workspace_intervals, self._epsabs, self._epsrel,
                self._epsilon, self._sigmas, kwargs.get('llzero'))


def xti_response():
    """second instrument: coarser grid, narrower band, smaller area"""
    m, edges, channels, ch_edges = syn.nicer_like_response(n_in=700, chan_lo=40, chan_hi=150)
    return np.ascontiguousarray(0.35 * m), edges, channels, ch_edges


def build_two_signals(counts_n, counts_x):
    interstellar = Interstellar(bounds=(0.0, 10.0))
    mN, eN, chN, ceN = syn.nicer_like_response()
    dataN = xpsi.Data(counts_n, channels=chN, phases=np.linspace(0.0, 1.0, 33), first=0, last=mN.shape[0] - 1,
                      exposure_time=syn.M2_EXPOSURE)
    sigN = mg.CustomSignal(data=dataN, instrument=xpsi.Instrument(mN, eN, chN, ceN), interstellar=interstellar,
                           cache=True, workspace_intervals=1000, epsrel=1.0e-8, epsilon=1.0e-3, sigmas=10.0,
                           prefix='N')
    mX, eX, chX, ceX = xti_response()
    dataX = xpsi.Data(counts_x, channels=chX, phases=np.linspace(0.0, 1.0, XTI_BINS + 1), first=0,
                      last=mX.shape[0] - 1, exposure_time=XTI_EXPOSURE)
    sigX = ShiftedSignal(data=dataX, instrument=xpsi.Instrument(mX, eX, chX, ceX), interstellar=interstellar,
                           cache=True, workspace_intervals=1000, epsrel=1.0e-8, epsilon=1.0e-3, sigmas=10.0,
                           bounds=dict(phase_shift=(-0.2, 0.2)), prefix='X')
    b = syn.M2_BOUNDS
    spacetime = xpsi.Spacetime(dict(mass=tuple(b[0]), radius=tuple(b[1]), distance=tuple(b[2]),
                                    cos_inclination=tuple(b[3])), values=dict(frequency=syn.M2_FREQUENCY))
    bounds = dict(super_colatitude=(None, None), super_radius=(None, None), phase_shift=(-0.25, 0.75),
                  super_temperature=(5.1, 6.8))
    primary = xpsi.HotRegion(bounds=bounds, values={}, symmetry=True, omit=False, cede=False, concentric=False,
                             sqrt_num_cells=32, min_sqrt_num_cells=10, max_sqrt_num_cells=64, num_leaves=100,
                             num_rays=200, atm_ext="Num4D", image_order_limit=3, prefix='p')

    class derive(xpsi.Derive):
        def __init__(self):
            pass

        def __call__(self, boundto, caller=None):
            return primary['super_temperature'] - syn.M2_SECONDARY_DT
    bounds = dict(bounds)
    bounds['super_temperature'] = None
    secondary = xpsi.HotRegion(bounds=bounds, values={'super_temperature': derive()}, symmetry=True, omit=False,
                               cede=False, concentric=False, sqrt_num_cells=32, min_sqrt_num_cells=10,
                               max_sqrt_num_cells=100, num_leaves=100, num_rays=200, is_antiphased=True,
                               atm_ext="Num4D", image_order_limit=3, prefix='s')
    hot = xpsi.HotRegions((primary, secondary))

    class Photosphere(xpsi.Photosphere):
        @xpsi.Photosphere.hot_atmosphere.setter
        def hot_atmosphere(self, table):
            self._hot_atmosphere = table
    photosphere = Photosphere(hot=hot, elsewhere=None, values=dict(mode_frequency=spacetime['frequency']))
    photosphere.hot_atmosphere = syn.nsx_like_table()
    star = xpsi.Star(spacetime=spacetime, photospheres=photosphere)
    like = xpsi.Likelihood(star=star, signals=[[sigN, sigX]], num_energies=128, threads=1,
                           externally_updated=False, prior=mg.FlatPrior())
    return like, sigN, sigX


def expected_of(signal, exposure, n_bins, background_rate):
    from xpsi.tools import phase_integrator
    phases = np.linspace(0.0, 1.0, n_bins + 1)
    out = np.zeros((signal.signals[0].shape[0], n_bins))
    for comp, sph, shift in zip(signal.signals, signal.phases, signal.shifts):
        out += phase_integrator(exposure, phases, comp, sph, shift)
    return out + background_rate * exposure / n_bins


def two_signals():
    mN = syn.nicer_like_response()[0]
    mX = xti_response()[0]
    like, sigN, sigX = build_two_signals(np.ones((mN.shape[0], 32)), np.ones((mX.shape[0], XTI_BINS)))
    names = list(like.names)
    print("two-signal parameter order:", names)
    vals = dict(zip(syn.M2_NAMES, syn.M2_TRUE))
    vals.update({"N__column_density": 0.6, "column_density": 0.6, "X__phase_shift": 0.03})
    theta = [vals[n] if n in vals else vals[n.split("__")[-1]] for n in names]
    like(theta, force=True)
    rng = np.random.default_rng(11)
    counts_n = rng.poisson(expected_of(sigN, syn.M2_EXPOSURE, 32, syn.M2_BACKGROUND_RATE)).astype(np.double)
    counts_x = rng.poisson(expected_of(sigX, XTI_EXPOSURE, XTI_BINS, 2.0e-3)).astype(np.double)
    print("synthetic data: %.4e + %.4e counts" % (counts_n.sum(), counts_x.sum()))
    like, sigN, sigX = build_two_signals(counts_n, counts_x)
    out = {"names": np.asarray(names), "counts_N": counts_n, "counts_X": counts_x,
           "energies": np.asarray(sigN.energies), "xti_exposure": np.asarray(XTI_EXPOSURE),
           "xti_bins": np.asarray(XTI_BINS), "attenuation_N_unit": np.exp(-0.3 * sigN._energy_mids ** -2.5),
           "attenuation_X_unit": np.exp(-0.3 * sigX._energy_mids ** -2.5)}
    assert np.array_equal(sigN.energies, sigX.energies)
    thetas = [list(theta), list(theta), list(theta)]
    thetas[1][names.index("p__super_colatitude")] += 0.01
    thetas[1][names.index("X__phase_shift")] = -0.07
    thetas[2][names.index("mass")] = 1.55
    thetas[2][names.index("s__super_radius")] = 0.3
    thetas[2][names.index("X__phase_shift")] = 0.11
    thetas[2][names.index([n for n in names if n.endswith("column_density")][0])] = 1.3
    out["theta"] = np.asarray(thetas)
    lnL, lN, lX, eN, eX = [], [], [], [], []
    for th in thetas:
        v = like(th, force=True)
        lnL.append(v); lN.append(sigN.loglikelihood); lX.append(sigX.loglikelihood)
        eN.append(np.array(sigN.expected_counts)); eX.append(np.array(sigX.expected_counts))
        print("lnL %.10f = %.10f + %.10f" % (v, lN[-1], lX[-1]))
    out.update(lnL_total=np.asarray(lnL), lnL_N=np.asarray(lN), lnL_X=np.asarray(lX),
               expected_N=np.asarray(eN), expected_X=np.asarray(eX))
    np.savez_compressed(os.path.join(HERE, "multi_signal.npz"), **out)


def build_everywhere(counts):
    interstellar = Interstellar(bounds=(0.0, 10.0))
    m, e, ch, ce = syn.nicer_like_response()
    data = xpsi.Data(counts, channels=ch, phases=np.array([0.0, 1.0]), first=0, last=m.shape[0] - 1,
                     exposure_time=EV_EXPOSURE)
    signal = mg.CustomSignal(data=data, instrument=xpsi.Instrument(m, e, ch, ce), interstellar=interstellar,
                             cache=True, workspace_intervals=1000, epsrel=1.0e-8, epsilon=1.0e-3, sigmas=10.0)
    st = xpsi.Spacetime(dict(mass=(1.0, 2.0), radius=(10.0, 14.0), distance=(0.1, 2.5), cos_inclination=(0.05, 0.95)),
                        values=dict(frequency=300.0))
    everywhere = xpsi.Everywhere(time_invariant=True, bounds=dict(temperature=(5.5, 6.6)), values={},
                                 sqrt_num_cells=24, num_rays=300, atm_ext="Num4D", image_order_limit=3)

    class Photosphere(xpsi.Photosphere):
        @xpsi.Photosphere.everywhere_atmosphere.setter
        def everywhere_atmosphere(self, table):
            self._everywhere_atmosphere = table
    ph = Photosphere(hot=None, elsewhere=None, everywhere=everywhere, values=dict(mode_frequency=300.0))
    ph.everywhere_atmosphere = syn.nsx_like_table()
    star = xpsi.Star(spacetime=st, photospheres=ph)
    like = xpsi.Likelihood(star=star, signals=signal, num_energies=128, threads=1, externally_updated=False,
                           prior=mg.FlatPrior())
    return like, signal, ph


def everywhere():
    n_chan = syn.nicer_like_response()[0].shape[0]
    like, signal, ph = build_everywhere(np.ones((n_chan, 1)))
    names = list(like.names)
    print("everywhere parameter order:", names)
    vals = dict(mass=1.4, radius=12.0, distance=1.5, cos_inclination=0.4, temperature=6.45, column_density=0.5)
    theta = [vals[n] for n in names]
    like(theta, force=True)
    # phase-averaged data: counts = T * (rate + background)
    expected = EV_EXPOSURE * (signal.signals[0][:, 0:1] + 5.0e-3)
    counts = np.random.default_rng(12).poisson(expected).astype(np.double)
    print("everywhere synthetic data: %.4e counts" % counts.sum())
    like, signal, ph = build_everywhere(counts)
    thetas = [list(theta), list(theta), list(theta)]
    thetas[1][names.index("temperature")] = 6.452
    thetas[1][names.index("radius")] = 12.004
    thetas[2][names.index("mass")] = 1.405
    thetas[2][names.index("cos_inclination")] = 0.41
    thetas[2][names.index("column_density")] = 0.52
    thetas[2][names.index("temperature")] = 6.4497
    out = {"names": np.asarray(names), "counts": counts, "theta": np.asarray(thetas),
           "energies": np.asarray(signal.energies), "exposure": np.asarray(EV_EXPOSURE),
           "attenuation_unit": np.exp(-0.3 * signal._energy_mids ** -2.5)}
    lnL, spec, folded, exp_ = [], [], [], []
    for th in thetas:
        v = like(th, force=True)
        lnL.append(v)
        spec.append(np.array(ph.signal[0][0][:, 0]))
        folded.append(np.array(signal.signals[0][:, 0]))
        exp_.append(np.array(signal.expected_counts))
        print("everywhere lnL %.10f" % v)
    out.update(lnL_total=np.asarray(lnL), spectrum=np.asarray(spec), folded=np.asarray(folded),
               expected=np.asarray(exp_))
    np.savez_compressed(os.path.join(HERE, "everywhere.npz"), **out)


if __name__ == "__main__":
    if "--everywhere-only" not in sys.argv:
        two_signals()
    everywhere()
    for f in ("multi_signal.npz", "everywhere.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
