#!/usr/bin/env python
"""Golden vectors for config 3 (SURVEY.md s8d "M3"): complex hot regions with the NSX-shaped atmosphere,
recorded from the reference build.  Primary = CST (superseding region with a concentric omission hole),
secondary = PDT (superseding + non-concentric ceding member).  Three integrator calls per evaluation
(primary super, secondary super, secondary cede); the two members of the secondary are summed after
energy integration (xpsi/Signal.py:419-429).

Usage: python oracle/build_ref.py && python tests/golden/make_golden_m3.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

xpsi, syn = mg.xpsi, mg.syn


def build_m3(rec, counts):
    matrix, edges, channels, ch_edges = syn.nicer_like_response()
    n_chan = matrix.shape[0]
    data = xpsi.Data(counts, channels=channels, phases=np.linspace(0.0, 1.0, 33), first=0, last=n_chan - 1,
                     exposure_time=syn.M2_EXPOSURE)
    instrument = xpsi.Instrument(matrix, edges, channels, ch_edges)
    signal = mg.CustomSignal(data=data, instrument=instrument, interstellar=None, cache=True,
                             workspace_intervals=1000, epsrel=1.0e-8, epsilon=1.0e-3, sigmas=10.0)
    b = syn.M2_BOUNDS
    spacetime = xpsi.Spacetime(dict(mass=tuple(b[0]), radius=tuple(b[1]), distance=tuple(b[2]),
                                    cos_inclination=tuple(b[3])), values=dict(frequency=syn.M2_FREQUENCY))
    bounds = dict(super_colatitude=(None, None), super_radius=(None, None), phase_shift=(-0.25, 0.75),
                  super_temperature=(5.1, 6.8), omit_radius=(None, None))
    primary = xpsi.HotRegion(bounds=bounds, values={}, symmetry=True, omit=True, cede=False, concentric=True,
                             sqrt_num_cells=32, min_sqrt_num_cells=10, max_sqrt_num_cells=80, num_leaves=64,
                             num_rays=512, atm_ext="Num4D", image_order_limit=3, prefix='p')
    bounds = dict(super_colatitude=(None, None), super_radius=(None, None), phase_shift=(-0.25, 0.75),
                  super_temperature=(5.1, 6.8), cede_colatitude=(None, None), cede_radius=(None, None),
                  cede_azimuth=(None, None), cede_temperature=(5.1, 6.8))
    secondary = xpsi.HotRegion(bounds=bounds, values={}, symmetry=True, omit=False, cede=True, concentric=False,
                               sqrt_num_cells=32, min_sqrt_num_cells=10, max_sqrt_num_cells=80, num_leaves=64,
                               num_rays=512, is_antiphased=True, atm_ext="Num4D", image_order_limit=3, prefix='s')
    hot = xpsi.HotRegions((primary, secondary))

    class Photosphere(xpsi.Photosphere):
        @xpsi.Photosphere.hot_atmosphere.setter
        def hot_atmosphere(self, table):
            self._hot_atmosphere = table
    photosphere = Photosphere(hot=hot, elsewhere=None, values=dict(mode_frequency=spacetime['frequency']))
    photosphere.hot_atmosphere = syn.nsx_like_table()
    star = xpsi.Star(spacetime=spacetime, photospheres=photosphere)
    like = xpsi.Likelihood(star=star, signals=signal, num_energies=128, threads=1, externally_updated=False,
                           prior=mg.FlatPrior())
    for h in (primary, secondary):
        rec.wrap_integrator(h)
    return like, signal


def main():
    rec = mg.Recorder()
    rec.wrap_signal_module()
    n_chan = syn.nicer_like_response()[0].shape[0]
    like, signal = build_m3(rec, np.ones((n_chan, 32)))
    print("M3 parameter order:", like.names)
    vals = {"mass": 1.5, "radius": 12.5, "distance": 1.5, "cos_inclination": 0.45,
            "p__phase_shift": 0.05, "p__super_colatitude": 0.9, "p__super_radius": 0.35, "p__omit_radius": 0.18,
            "p__super_temperature": 6.5,
            "s__phase_shift": -0.02, "s__super_colatitude": 2.1, "s__super_radius": 0.15,
            "s__cede_colatitude": 2.0, "s__cede_radius": 0.32, "s__cede_azimuth": 0.12,
            "s__super_temperature": 6.55, "s__cede_temperature": 6.4}
    theta = [vals[n] for n in like.names]
    like(theta, force=True)
    from xpsi.tools import phase_integrator
    phases = np.linspace(0.0, 1.0, 33)
    expected = np.zeros((n_chan, 32))
    for comp, sph, shift in zip(signal.signals, signal.phases, signal.shifts):
        expected += phase_integrator(syn.M2_EXPOSURE, phases, comp, sph, shift)
    expected += syn.M2_BACKGROUND_RATE * syn.M2_EXPOSURE / 32.0
    counts = np.random.default_rng(3).poisson(expected).astype(np.double)
    print("M3 synthetic data: total counts %.4e" % counts.sum())
    like, signal = build_m3(rec, counts)
    theta[like.names.index("s__cede_azimuth")] += 0.02
    out = {"counts": counts, "names": np.asarray(like.names)}
    lnL = mg.record_eval(like, signal, None, rec, theta, out, "", slim=True, keep_eint=False)
    print("M3 lnL = %.10f, integrator calls %d" % (lnL, int(out["n_members"])))
    out["d_sq"] = np.asarray(like.star.spacetime.d_sq)
    out["member_component"] = np.asarray([0, 1, 1])
    np.savez_compressed(os.path.join(HERE, "m3_cst_pdt.npz"), **out)
    rec.restore()
    print("m3_cst_pdt.npz", os.path.getsize(os.path.join(HERE, "m3_cst_pdt.npz")) // 1024, "KiB")
    for m in range(int(out["n_members"])):
        a = out["int%d_cellArea" % m]
        print(" member", m, "mesh", a.shape, "radiating cells", int((a > 0).sum()), "partial rows",
              int(((a > 0).sum(axis=1) < a.shape[1]).sum()))


if __name__ == "__main__":
    main()
