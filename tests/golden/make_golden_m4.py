#!/usr/bin/env python
"""Golden vectors for config 4 (SURVEY.md s8d "M4"), recorded from the reference build:

  m4_elsewhere.npz   ST-U NSX hot regions + Elsewhere (Num4D, 32x32 mesh, 400 rays) + interstellar
                     attenuation: the time-invariant integrator call, both hot-region integrator calls
                     with the elsewhere *correction* active, attenuated/folded signals and lnL
  m4_everywhere.npz  Everywhere(time_invariant=True) spectra: blackbody and Num4D

Usage: python oracle/build_ref.py && python tests/golden/make_golden_m4.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference, wraps nothing yet)

xpsi, syn = mg.xpsi, mg.syn

TINV_ARGS = ("numThreads", "R", "omega", "r_s", "inclination", "sqrt_numPix", "cellArea",
             "radialCoords_of_parallels", "r_s_over_r", "theta", "phi", "srcCellParams", "numRays",
             "deflection", "cos_alpha", "maxDeflection", "cos_gammaArray", "energies", "atmosphere",
             "atm_ext", "image_order_limit")


def wrap_tinv(modname, store):
    mod = sys.modules[modname]
    inner = mod._integrator

    def wrapped(*args):
        out = inner(*args)
        store.append((args, out))
        return out
    mod._integrator = wrapped
    return mod, inner


def pack_tinv(prefix, call, out):
    args, res = call
    for name, val in zip(TINV_ARGS, args):
        if name == "atmosphere" or val is None:
            continue
        out[prefix + name] = np.asarray(val)
    out[prefix + "status"] = np.asarray(res[0])
    out[prefix + "flux"] = np.asarray(res[1])


class Interstellar(xpsi.Interstellar):
    """attenuation = exp(-0.3 E^-2.5) ** N_H  (shape of CustomInterstellar.py:38-51, closed form)"""

    def __init__(self, bounds):
        nh = xpsi.Parameter('column_density', strict_bounds=(0.0, 10.0), bounds=bounds, doc='N_H',
                            symbol='N_H', value=None)
        super(Interstellar, self).__init__(nh)

    def attenuation(self, energies):
        return np.exp(-0.3 * energies ** -2.5) ** self['column_density']


def build_m4(rec, counts, tinv_store):
    matrix, edges, channels, ch_edges = syn.nicer_like_response()
    n_chan = matrix.shape[0]
    data = xpsi.Data(counts, channels=channels, phases=np.linspace(0.0, 1.0, 33), first=0, last=n_chan - 1,
                     exposure_time=syn.M2_EXPOSURE)
    instrument = xpsi.Instrument(matrix, edges, channels, ch_edges)
    interstellar = Interstellar(bounds=(0.0, 10.0))
    signal = mg.CustomSignal(data=data, instrument=instrument, interstellar=interstellar, cache=True,
                             workspace_intervals=1000, epsrel=1.0e-8, epsilon=1.0e-3, sigmas=10.0)
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
    elsewhere = xpsi.Elsewhere(bounds=dict(elsewhere_temperature=(5.2, 6.5)), values={}, sqrt_num_cells=32,
                               num_rays=400, atm_ext="Num4D", image_order_limit=3)

    class Photosphere(xpsi.Photosphere):
        @xpsi.Photosphere.hot_atmosphere.setter
        def hot_atmosphere(self, table):
            self._hot_atmosphere = table

        @xpsi.Photosphere.elsewhere_atmosphere.setter
        def elsewhere_atmosphere(self, table):
            self._elsewhere_atmosphere = table
    photosphere = Photosphere(hot=hot, elsewhere=elsewhere, values=dict(mode_frequency=spacetime['frequency']))
    table = syn.nsx_like_table()
    photosphere.hot_atmosphere = table
    photosphere.elsewhere_atmosphere = table
    star = xpsi.Star(spacetime=spacetime, photospheres=photosphere)
    like = xpsi.Likelihood(star=star, signals=signal, num_energies=128, threads=1, externally_updated=False,
                           prior=mg.FlatPrior())
    for h in (primary, secondary):
        rec.wrap_integrator(h)
    return like, signal, instrument


def main():
    rec = mg.Recorder()
    rec.wrap_signal_module()
    tinv = []
    mod, inner = wrap_tinv("xpsi.Elsewhere", tinv)
    n_chan = syn.nicer_like_response()[0].shape[0]
    like, signal, instrument = build_m4(rec, np.ones((n_chan, 32)), tinv)
    print("M4 parameter order:", like.names)
    # theta = M2_TRUE + elsewhere temperature + column density, in the likelihood's own order
    vals = dict(zip(syn.M2_NAMES, syn.M2_TRUE))
    vals.update({"elsewhere_temperature": 6.2, "column_density": 0.8})
    theta = [vals[n] for n in like.names]
    # synthetic data of this model: Poisson(default_rng(4)) of its own expected counts
    like(theta, force=True)
    from xpsi.tools import phase_integrator
    phases = np.linspace(0.0, 1.0, 33)
    expected = np.zeros((n_chan, 32))
    for comp, sph, shift in zip(signal.signals, signal.phases, signal.shifts):
        expected += phase_integrator(syn.M2_EXPOSURE, phases, comp, sph, shift)
    expected += syn.M2_BACKGROUND_RATE * syn.M2_EXPOSURE / 32.0
    counts = np.random.default_rng(4).poisson(expected).astype(np.double)
    print("M4 synthetic data: total counts %.4e" % counts.sum())
    mod._integrator = inner
    del tinv[:]
    mod, inner = wrap_tinv("xpsi.Elsewhere", tinv)
    like, signal, instrument = build_m4(rec, counts, tinv)
    theta[like.names.index("p__super_colatitude")] += 0.01     # evaluate slightly off the truth
    out = {"counts": counts}
    rec.clear()
    del tinv[:]
    lnL = like(theta, force=True)
    print("M4 lnL = %.10f" % lnL)
    out["theta"] = np.asarray(theta)
    out["names"] = np.asarray(like.names)
    out["lnL_total"] = np.asarray(lnL)
    out["d_sq"] = np.asarray(like.star.spacetime.d_sq)
    out["column_density"] = np.asarray(vals["column_density"])
    for m, call in enumerate(rec.calls["integrate"]):
        mg.pack_integrate("int%d_" % m, call, out)
        out["int%d_else_atm_ext" % m] = np.asarray(call[0][25])
    pack_tinv("else_", tinv[0], out)
    out["energy_mids"] = np.asarray(signal._energy_mids)
    out["attenuation"] = signal._interstellar.attenuation(signal._energy_mids)
    mg.pack_marginal("marg_", rec.calls["marginal"][0], out)
    mod._integrator = inner
    np.savez_compressed(os.path.join(HERE, "m4_elsewhere.npz"), **out)

    # ---- Everywhere(time_invariant=True): BB and Num4D spectra through the same integrator ----
    tinv2 = []
    ev = {}
    for tag, atm in (("bb", "BB"), ("num4d", "Num4D")):
        st = xpsi.Spacetime(dict(mass=(1.0, 2.0), radius=(10.0, 14.0), distance=(0.1, 2.5), cos_inclination=(0.05, 0.95)),
                            values=dict(frequency=300.0))
        everywhere = xpsi.Everywhere(time_invariant=True, bounds=dict(temperature=(5.5, 6.6)), values={},
                                     sqrt_num_cells=24, num_rays=300, atm_ext=atm, image_order_limit=3)

        class Photosphere(xpsi.Photosphere):
            @xpsi.Photosphere.everywhere_atmosphere.setter
            def everywhere_atmosphere(self, table):
                self._everywhere_atmosphere = table
        inner_ev = everywhere._integrator

        def wrapped_ev(*args, _inner=inner_ev):
            res = _inner(*args)
            tinv2.append((args, res))
            return res
        everywhere._integrator = wrapped_ev
        ph = Photosphere(hot=None, elsewhere=None, everywhere=everywhere, values=dict(mode_frequency=300.0))
        if atm == "Num4D":
            ph.everywhere_atmosphere = syn.nsx_like_table()
        star = xpsi.Star(spacetime=st, photospheres=ph)
        star([1.4, 12.0, 0.3, 0.4, 6.3])
        star.update()
        energies = np.logspace(np.log10(0.2), np.log10(7.7), 128)
        ph.integrate(energies, 1)
        pack_tinv(tag + "_", tinv2[-1], ev)
    np.savez_compressed(os.path.join(HERE, "m4_everywhere.npz"), **ev)
    rec.restore()
    for f in ("m4_elsewhere.npz", "m4_everywhere.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
