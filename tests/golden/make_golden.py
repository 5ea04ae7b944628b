#!/usr/bin/env python
"""Generate golden input/output vectors for the likelihood hot path by running
the reference build in oracle/_ref (unmodified X-PSI 3.3.0 sources, GSL-subset
shim) and recording every array that crosses a drop-in boundary of SURVEY.md
section 8b:

  * ``cellmesh.integrator_for_azimuthal_invariance.integrate`` args -> flux
  * ``tools.energy_integrator`` args -> integrated signal
  * ``Instrument.__call__``  (matrix, signal) -> folded signal
  * ``likelihoods.eval_marginal_likelihood`` args -> (lnL, expected counts,
    ML backgrounds)

Fixtures written next to this file:

  c1_st_bb.npz    examples_fast ST blackbody at the reference's known-answer
                  point (xpsi/tests/test_likelihood.py:134, lnL=-47881.278...)
  m2_stu_nsx.npz  ST-U, synthetic NSX-shaped table, NICER-like response,
                  synthetic Poisson data (SURVEY.md s8d "M2"), 2 theta

Run here (needs /root/reference once, to build oracle/_ref):
    python oracle/build_ref.py && python tests/golden/make_golden.py
"""
import math
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
sys.path.append(os.path.join(ref_env.REF_ROOT, "examples_fast", "Modules"))
from CustomSignal import CustomSignal  # noqa: E402
import CustomSignal as CustomSignal_module  # noqa: E402

INTEGRATE_ARGS = ("numThreads", "R", "omega", "r_s", "inclination", "cellArea",
                  "radialCoords_of_parallels", "r_s_over_r", "theta", "phi",
                  "srcCellParams", "CELL_RADIATES", "correction_srcCellParams",
                  "numRays", "deflection", "cos_alpha", "lag", "maxDeflection",
                  "cos_gammaArray", "energies", "leaves", "phases",
                  "hot_atmosphere", "elsewhere_atmosphere", "hot_atm_ext",
                  "else_atm_ext", "beam_opt", "image_order_limit")


class Recorder:
    """Wrap the reference's boundary callables and keep (args, result)."""

    def __init__(self):
        self.calls = {"integrate": [], "energy_integrator": [], "marginal": []}

    def wrap_integrator(self, hot):
        inner = hot._integrator

        def wrapped(*args, **kwargs):
            out = inner(*args, **kwargs)
            # copy: Photosphere.integrate later adds the elsewhere spectrum to this array in place
            # (xpsi/Photosphere.py:589-592)
            kept = (out[0], None if out[1] is None else np.array(out[1]))
            self.calls["integrate"].append((args, kwargs, kept))
            return out
        hot._integrator = wrapped

    def wrap_signal_module(self):
        mod = sys.modules["xpsi.Signal"]
        inner = mod.energy_integrator

        def wrapped(*args):
            out = inner(*args)
            self.calls["energy_integrator"].append(([np.array(a) if isinstance(a, np.ndarray) else a for a in args], out.copy()))
            return out
        mod.energy_integrator = wrapped
        self._restore_signal = (mod, inner)

        inner_m = CustomSignal_module.eval_marginal_likelihood

        def wrapped_m(*args):
            a = [np.array(x) if isinstance(x, np.ndarray) else x for x in args]
            a[3] = tuple(np.array(x) for x in args[3])
            a[4] = tuple(np.array(x) for x in args[4])
            out = inner_m(*args)
            self.calls["marginal"].append((a, out))
            return out
        CustomSignal_module.eval_marginal_likelihood = wrapped_m
        self._restore_marg = inner_m

    def restore(self):
        mod, inner = self._restore_signal
        mod.energy_integrator = inner
        CustomSignal_module.eval_marginal_likelihood = self._restore_marg

    def clear(self):
        for v in self.calls.values():
            del v[:]


def pack_integrate(prefix, call, out):
    args, kwargs, res = call
    for name, val in zip(INTEGRATE_ARGS, args):
        if name in ("hot_atmosphere", "elsewhere_atmosphere"):
            continue
        if val is None:
            continue
        out[prefix + name] = np.asarray(val)
    out[prefix + "status"] = np.asarray(res[0])
    out[prefix + "flux"] = np.asarray(res[1])


def pack_marginal(prefix, call, out):
    a, res = call
    names = ("exposure_time", "phases", "counts", "components", "component_phases",
             "phase_shifts", "precomp", "support", "workspace_intervals", "epsabs",
             "epsrel", "epsilon", "sigmas", "llzero")
    for name, val in zip(names, a):
        if name in ("components", "component_phases"):
            for i, c in enumerate(val):
                out["%s%s_%d" % (prefix, name, i)] = np.asarray(c)
            out[prefix + "n_components"] = np.asarray(len(val))
        else:
            out[prefix + name] = np.asarray(val)
    out[prefix + "lnL"] = np.asarray(res[0])
    out[prefix + "expected_counts"] = np.asarray(res[1])
    out[prefix + "mcl_background"] = np.asarray(res[2])
    out[prefix + "mcl_background_given_support"] = np.asarray(res[3])


class FlatPrior(xpsi.Prior):
    __derived_names__ = []

    def __call__(self, p=None):
        return 0.0

    def inverse_sample(self, hypercube=None):
        return None

    def transform(self, p, **kwargs):
        return p


# --------------------------------------------------------------------------
def build_c1(rec):
    counts = np.loadtxt(os.path.join(ref_env.REF_ROOT, "examples_fast", "Data",
                                     "xpsi_good_realisation.dat"), dtype=np.double)
    data = xpsi.Data(counts, channels=np.arange(10, 301), phases=np.linspace(0.0, 1.0, 33),
                     first=0, last=290, exposure_time=1000.0)
    matrix, edges, channels, ch_edges = syn.c1_response()
    instrument = xpsi.Instrument(matrix, edges, channels, ch_edges)
    signal = CustomSignal(data=data, instrument=instrument, interstellar=None, cache=True,
                          workspace_intervals=1000, epsrel=1.0e-8, epsilon=1.0e-3, sigmas=10.0)
    spacetime = xpsi.Spacetime(dict(distance=(0.5, 2), mass=(1.0, 1.6), radius=(10, 13),
                                    cos_inclination=(0, 1)), values=dict(frequency=314.0))
    bounds = dict(super_colatitude=(0.001, math.pi / 2 - 0.001),
                  super_radius=(0.001, math.pi / 2 - 0.001),
                  phase_shift=(-0.25, 0.75), super_temperature=(6., 7.))
    hot = xpsi.HotRegion(bounds=bounds, values={}, symmetry=True, omit=False, cede=False,
                         concentric=False, sqrt_num_cells=32, min_sqrt_num_cells=16,
                         max_sqrt_num_cells=64, num_leaves=64, num_rays=512,
                         is_antiphased=True, image_order_limit=3, prefix='hot')
    photosphere = xpsi.Photosphere(hot=hot, elsewhere=None,
                                   values=dict(mode_frequency=spacetime['frequency']))
    star = xpsi.Star(spacetime=spacetime, photospheres=photosphere)
    like = xpsi.Likelihood(star=star, signals=signal, num_energies=128, threads=1,
                           externally_updated=False, prior=FlatPrior())
    rec.wrap_integrator(hot)
    return like, signal, instrument, [hot]


def build_m2(rec, counts):
    matrix, edges, channels, ch_edges = syn.nicer_like_response()
    n_chan = matrix.shape[0]
    data = xpsi.Data(counts, channels=channels, phases=np.linspace(0.0, 1.0, 33),
                     first=0, last=n_chan - 1, exposure_time=syn.M2_EXPOSURE)
    instrument = xpsi.Instrument(matrix, edges, channels, ch_edges)
    signal = CustomSignal(data=data, instrument=instrument, interstellar=None, cache=True,
                          workspace_intervals=1000, epsrel=1.0e-8, epsilon=1.0e-3, sigmas=10.0)
    b = syn.M2_BOUNDS
    spacetime = xpsi.Spacetime(dict(mass=tuple(b[0]), radius=tuple(b[1]), distance=tuple(b[2]),
                                    cos_inclination=tuple(b[3])),
                               values=dict(frequency=syn.M2_FREQUENCY))
    bounds = dict(super_colatitude=(None, None), super_radius=(None, None),
                  phase_shift=(-0.25, 0.75), super_temperature=(5.1, 6.8))
    primary = xpsi.HotRegion(bounds=bounds, values={}, symmetry=True, omit=False, cede=False,
                             concentric=False, sqrt_num_cells=32, min_sqrt_num_cells=10,
                             max_sqrt_num_cells=64, num_leaves=100, num_rays=200,
                             atm_ext="Num4D", image_order_limit=3, prefix='p')

    class derive(xpsi.Derive):
        def __init__(self):
            pass

        def __call__(self, boundto, caller=None):
            return primary['super_temperature'] - syn.M2_SECONDARY_DT

    bounds = dict(bounds)
    bounds['super_temperature'] = None
    secondary = xpsi.HotRegion(bounds=bounds, values={'super_temperature': derive()},
                               symmetry=True, omit=False, cede=False, concentric=False,
                               sqrt_num_cells=32, min_sqrt_num_cells=10, max_sqrt_num_cells=100,
                               num_leaves=100, num_rays=200, is_antiphased=True,
                               atm_ext="Num4D", image_order_limit=3, prefix='s')
    hot = xpsi.HotRegions((primary, secondary))

    class Photosphere(xpsi.Photosphere):
        @xpsi.Photosphere.hot_atmosphere.setter
        def hot_atmosphere(self, table):
            self._hot_atmosphere = table

    photosphere = Photosphere(hot=hot, elsewhere=None,
                              values=dict(mode_frequency=spacetime['frequency']))
    photosphere.hot_atmosphere = syn.nsx_like_table()
    star = xpsi.Star(spacetime=spacetime, photospheres=photosphere)
    like = xpsi.Likelihood(star=star, signals=signal, num_energies=128, threads=1,
                           externally_updated=False, prior=FlatPrior())
    for h in (primary, secondary):
        rec.wrap_integrator(h)
    return like, signal, instrument, [primary, secondary]


def record_eval(like, signal, instrument, rec, theta, out, prefix, slim=False, keep_eint=True):
    rec.clear()
    lnL = like(list(theta), force=True)
    out[prefix + "theta"] = np.asarray(theta, dtype=np.double)
    out[prefix + "lnL_total"] = np.asarray(lnL)
    n_int = len(rec.calls["integrate"])
    out[prefix + "n_members"] = np.asarray(n_int)
    for m, call in enumerate(rec.calls["integrate"]):
        pack_integrate("%sint%d_" % (prefix, m), call, out)
    # energy_integrator input m is integrate-flux m / d_sq (Likelihood.py:361-364)
    # and the folded signals are the marginal-likelihood components, so neither
    # is stored twice; ``slim`` keeps one integrated signal only.
    for m, (a, res) in enumerate(rec.calls["energy_integrator"]):
        if m == 0:
            out[prefix + "eint_log10_energies"] = a[2]
            out[prefix + "eint_log10_edges"] = a[3]
        if (slim and m > 0) or not keep_eint:
            continue
        out["%seint%d_out" % (prefix, m)] = res
    pack_marginal(prefix + "marg_", rec.calls["marginal"][0], out)
    return lnL


def main():
    rec = Recorder()
    rec.wrap_signal_module()

    # ---------------- C1 ---------------------------------------------------
    like, signal, instrument, hots = build_c1(rec)
    out = {}
    p = [1.4, 10, 1., math.cos(60 * np.pi / 180), 0.0, 70 * np.pi / 180, 0.75, 6.8]
    lnL = record_eval(like, signal, instrument, rec, p, out, "")
    print("C1 lnL = %.12f (reference known answer -47881.27817666349, rel %.2e)"
          % (lnL, abs(lnL + 47881.27817666349) / 47881.27817666349))
    assert abs(lnL + 47881.27817666349) < 1e-5 * 47881.27817666349
    out["known_answer_lnL"] = np.asarray(-47881.27817666349)
    out["d_sq"] = np.asarray(like.star.spacetime.d_sq)
    np.savez_compressed(os.path.join(HERE, "c1_st_bb.npz"), **out)

    # ---------------- M2 ---------------------------------------------------
    matrix = syn.nicer_like_response()[0]
    n_chan = matrix.shape[0]
    like, signal, instrument, hots = build_m2(rec, np.ones((n_chan, 32)))
    scratch = {}
    record_eval(like, signal, instrument, rec, syn.M2_TRUE, scratch, "")
    # expected counts at theta_true -> synthetic Poisson data (default_rng(0))
    from xpsi.tools import phase_integrator
    phases = np.linspace(0.0, 1.0, 33)
    expected = np.zeros((n_chan, 32))
    for comp, sph, shift in zip(signal.signals, signal.phases, signal.shifts):
        expected += phase_integrator(syn.M2_EXPOSURE, phases, comp, sph, shift)
    expected += syn.M2_BACKGROUND_RATE * syn.M2_EXPOSURE / 32.0
    counts = np.random.default_rng(0).poisson(expected).astype(np.double)
    print("M2 synthetic data: total counts %.4e, max bin %d" % (counts.sum(), counts.max()))

    like, signal, instrument, hots = build_m2(rec, counts)
    out = {"counts": counts, "expected_true": expected}
    thetas = [syn.M2_TRUE, syn.m2_theta_batch(4)[1]]
    for t, th in enumerate(thetas):
        lnL = record_eval(like, signal, instrument, rec, th, out, "t%d_" % t, slim=True,
                          keep_eint=(t == 0))
        print("M2 theta %d lnL = %.10f" % (t, lnL))
        out["t%d_d_sq" % t] = np.asarray(like.star.spacetime.d_sq)
    out["n_theta"] = np.asarray(len(thetas))
    np.savez_compressed(os.path.join(HERE, "m2_stu_nsx.npz"), **out)
    rec.restore()
    for f in ("c1_st_bb.npz", "m2_stu_nsx.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
