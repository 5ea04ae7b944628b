#!/usr/bin/env python
"""Golden vectors for hot regions that cover a rotational pole (cellmesh/polar_mesh.pyx): the M2 (ST-U, Num4D)
model evaluated with a northern polar cap as primary and a southern polar cap as secondary.  Records the
embedded meshes / rays (integrator inputs), fluxes, folded signals and lnL from the reference build; the
data are a Poisson realisation of the model's own expected counts so the likelihood is well conditioned.
Only the mesh-level arrays needed by the parity test are kept (no integrated signals).

Usage: python oracle/build_ref.py && python tests/golden/make_golden_polar.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

xpsi, syn = mg.xpsi, mg.syn


def main():
    rec = mg.Recorder()
    rec.wrap_signal_module()
    n_chan = syn.nicer_like_response()[0].shape[0]
    like, signal, instrument, hot = mg.build_m2(rec, np.ones((n_chan, 32)))
    vals = {"mass": 1.45, "radius": 12.2, "distance": 1.2, "cos_inclination": 0.55,
            "p__phase_shift": 0.1, "p__super_colatitude": 0.2, "p__super_radius": 0.35, "p__super_temperature": 6.45,
            "s__phase_shift": -0.05, "s__super_colatitude": np.pi - 0.12, "s__super_radius": 0.3}
    theta = [vals[n] for n in like.names]
    like(theta, force=True)
    from xpsi.tools import phase_integrator
    phases = np.linspace(0.0, 1.0, 33)
    expected = np.zeros((n_chan, 32))
    for comp, sph, shift in zip(signal.signals, signal.phases, signal.shifts):
        expected += phase_integrator(syn.M2_EXPOSURE, phases, comp, sph, shift)
    expected += syn.M2_BACKGROUND_RATE * syn.M2_EXPOSURE / 32.0
    counts = np.random.default_rng(5).poisson(expected).astype(np.double)
    print("polar synthetic data: total counts %.4e" % counts.sum())
    rec.clear()
    like, signal, instrument, hot = mg.build_m2(rec, counts)
    out = {"counts": counts, "names": np.asarray(like.names)}
    lnL = mg.record_eval(like, signal, None, rec, theta, out, "", slim=True, keep_eint=False)
    out["d_sq"] = np.asarray(like.star.spacetime.d_sq)
    for k in [k for k in out if k.startswith("marg_expected") or k.startswith("marg_mcl")]:
        del out[k]
    np.savez_compressed(os.path.join(HERE, "m5_polar.npz"), **out)
    rec.restore()
    print("polar lnL = %.10f" % lnL)
    print("m5_polar.npz", os.path.getsize(os.path.join(HERE, "m5_polar.npz")) // 1024, "KiB")
    for m in range(int(out["n_members"])):
        a = out["int%d_cellArea" % m]
        print(" member", m, "mesh", a.shape, "radiating cells", int((a > 0).sum()), "theta range",
              float(out["int%d_theta" % m].min()), float(out["int%d_theta" % m].max()))


if __name__ == "__main__":
    main()
