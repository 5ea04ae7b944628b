#!/usr/bin/env python
"""Golden vectors for the stand-alone signal tools, recorded from the
reference build in oracle/_ref: tools.phase_integrator, phase_interpolator, energy_interpolator.
Inputs are the C1 fixture's folded signal / flux, so the shapes are the reference's own."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_env  # noqa: E402

xpsi = ref_env.import_reference()
from xpsi.tools import energy_interpolator, phase_integrator, phase_interpolator  # noqa: E402

c1 = np.load(os.path.join(HERE, "c1_st_bb.npz"))
pulse = np.ascontiguousarray(c1["marg_components_0"][::8])            # [37, 64] count-rate rows
sig_phases = c1["marg_component_phases_0"]
edges = c1["marg_phases"]
flux = np.ascontiguousarray(c1["int0_flux"][:, ::4])                  # [128, 16]
log10E = np.log10(c1["int0_energies"])
new_E = np.linspace(log10E[0], log10E[-1] + 0.05, 57)                   # a few points beyond the last energy
out = {"pulse": pulse, "sig_phases": sig_phases, "edges": edges, "flux": flux, "log10E": log10E, "new_E": new_E}
for shift in (0.0, 0.37, -0.2):
    tag = ("%+.2f" % shift).replace(".", "p").replace("+", "P").replace("-", "M")
    out["pint_" + tag] = phase_integrator(1000.0, edges, pulse, sig_phases, shift)
    out["pitp_" + tag] = phase_interpolator(np.linspace(0.0, 1.0, 41), sig_phases, pulse, shift)
out["shifts"] = np.array([0.0, 0.37, -0.2])
out["new_phases"] = np.linspace(0.0, 1.0, 41)
out["eitp"] = energy_interpolator(1, flux, log10E, new_E)
flux_neg = flux.copy(); flux_neg[5, 3] = 0.0                           # forces the linear (non-log) mode in one column
out["flux_neg"] = flux_neg
out["eitp_neg"] = energy_interpolator(1, flux_neg, log10E, new_E)
np.savez_compressed(os.path.join(HERE, "tools.npz"), **out)
print("tools.npz", os.path.getsize(os.path.join(HERE, "tools.npz")) // 1024, "KiB")

# ---- _poisson_likelihood_given_background (row f4) on the C1 folded signal ----
from xpsi.likelihoods._poisson_likelihood_given_background import poisson_likelihood_given_background  # noqa: E402
comp = c1["marg_components_0"]
bgr = np.abs(np.sin(np.arange(comp.shape[0] * 32, dtype=np.double)).reshape(comp.shape[0], 32)) * 0.05 + 0.01
lnL_bg, expec_bg = poisson_likelihood_given_background(1000.0, edges, c1["marg_counts"], (comp,), (sig_phases,),
                                                       np.array([0.13]), bgr, c1["marg_precomp"])
out = dict(np.load(os.path.join(HERE, "tools.npz")))
out.update({"plgb_background": bgr, "plgb_lnL": np.asarray(lnL_bg), "plgb_expected": np.asarray(expec_bg),
            "plgb_shift": np.asarray(0.13)})
np.savez_compressed(os.path.join(HERE, "tools.npz"), **out)
print("tools.npz (+given-background likelihood)", os.path.getsize(os.path.join(HERE, "tools.npz")) // 1024, "KiB")

# ---- tools.synthesise (row b / f4): expected counts and scales only, the Poisson draw is not recorded ----
from xpsi.tools.synthesise import synthesise_exposure, synthesise_given_total_count_number  # noqa: E402
bg_counts = bgr * 1000.0
e1, _, s1 = synthesise_exposure(1000.0, edges, (comp,), (sig_phases,), np.array([0.13]), 5.0e4, bg_counts, gsl_seed=1)
e2, _, s2a, s2b = synthesise_given_total_count_number(edges, 2.0e6, (comp,), (sig_phases,), np.array([0.13]), 5.0e4,
                                                     bg_counts, gsl_seed=1)
out = dict(np.load(os.path.join(HERE, "tools.npz")))
out.update({"syn_bg_counts": bg_counts, "syn_exposure_expected": np.asarray(e1), "syn_exposure_scale": np.asarray(s1),
            "syn_total_expected": np.asarray(e2), "syn_total_scales": np.array([s2a, s2b])})
np.savez_compressed(os.path.join(HERE, "tools.npz"), **out)
print("tools.npz (+synthesise)", os.path.getsize(os.path.join(HERE, "tools.npz")) // 1024, "KiB")

# ---- eval_marginal_likelihood with more than 32 data phase bins (64 and a ragged 45) --------------------
from xpsi.likelihoods.default_background_marginalisation import eval_marginal_likelihood, precomputation  # noqa: E402
from xpsi.tools import phase_integrator as _pint  # noqa: E402
out = dict(np.load(os.path.join(HERE, "tools.npz")))
rng = np.random.default_rng(11)
for nb in (64, 45):
    ph = np.linspace(0.0, 1.0, nb + 1)
    expec = _pint(1000.0, ph, comp, sig_phases, 0.13) + 2.0 * 1000.0 / nb
    cnts = rng.poisson(expec).astype(np.double)
    pre = precomputation(cnts.astype(np.int32))
    sup = -1.0 * np.ones((comp.shape[0], 2)); sup[:, 0] = 0.0
    res = eval_marginal_likelihood(1000.0, ph, cnts, (comp,), (sig_phases,), np.array([0.13]), pre, sup,
                                   1000, 0.0, 1.0e-8, 1.0e-3, 10.0, -1.0e90)
    out.update({"mbins%d_counts" % nb: cnts, "mbins%d_lnL" % nb: np.asarray(res[0]),
                "mbins%d_expected" % nb: np.asarray(res[1]), "mbins%d_bg" % nb: np.asarray(res[2])})
    print("marginal likelihood with %d bins: lnL = %.8f" % (nb, res[0]))
np.savez_compressed(os.path.join(HERE, "tools.npz"), **out)
print("tools.npz (+many-bin likelihood)", os.path.getsize(os.path.join(HERE, "tools.npz")) // 1024, "KiB")

# ---- time-invariant component (Everywhere(time_invariant=True)): one signal column, phase-averaged data ----
# compute_expected_counts.pyx:190-192 stores the rate in the first bin instead of integrating a spline
ev = np.load(os.path.join(HERE, "m4_everywhere.npz"))
from xpsi.tools import energy_integrator as _eint  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from xpsi_b200 import synthetic as _syn  # noqa: E402
_matrix, _edges = _syn.nicer_like_response()[:2]
spec = np.ascontiguousarray(ev["num4d_flux"].reshape(-1, 1) / (1.2 * 3.08567758e19) ** 2)
integ = _eint(1, spec, np.log10(ev["num4d_energies"]), np.log10(_edges))
folded = np.ascontiguousarray(np.dot(_matrix, integ))                    # [n_chan, 1]
T_ev = 2.0e4
one_bin = np.array([0.0, 1.0])
cnts = np.random.default_rng(21).poisson(folded * T_ev + 40.0).astype(np.double)     # no empty channel: with one bin the reference fails on zero counts (infinite quadrature bounds)
pre = precomputation(cnts.astype(np.int32))
sup = -1.0 * np.ones((folded.shape[0], 2)); sup[:, 0] = 0.0
res = eval_marginal_likelihood(T_ev, one_bin, cnts, (folded,), (np.array([0.0]),), np.array([0.0]), pre, sup,
                               1000, 0.0, 1.0e-8, 1.0e-3, 10.0, -1.0e90)
out = dict(np.load(os.path.join(HERE, "tools.npz")))
out.update({"tinv_folded": folded, "tinv_counts": cnts, "tinv_exposure": np.asarray(T_ev),
            "tinv_lnL": np.asarray(res[0]), "tinv_expected": np.asarray(res[1]), "tinv_bg": np.asarray(res[2])})
np.savez_compressed(os.path.join(HERE, "tools.npz"), **out)
print("time-invariant likelihood lnL = %.8f" % res[0])
print("tools.npz (+time-invariant likelihood)", os.path.getsize(os.path.join(HERE, "tools.npz")) // 1024, "KiB")
