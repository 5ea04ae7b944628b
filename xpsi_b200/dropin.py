"""Run-time rebinding of the reference's compiled seams to the GPU callables (INTEGRATION.md sections 1-3).

``install(xpsi)`` takes an *imported* X-PSI package object and rebinds, in place, exactly the names a maintainer
would edit by hand:

==========================================================  ==================================================
reference name (file:line)                                  bound to
==========================================================  ==================================================
``xpsi.cellmesh.integrator_for_azimuthal_invariance``       ``xpsi_b200.cellmesh.integrator_for_azimuthal_
``.integrate`` (HotRegion.py:551-572 imports it late)       invariance.integrate``
``xpsi.cellmesh.integrator.integrate`` (HotRegion.py:567,   ``xpsi_b200.cellmesh.integrator.integrate``
Everywhere.py:330)
``xpsi.cellmesh.integrator_for_time_invariance.integrate``  ``xpsi_b200.cellmesh.integrator_for_time_
(Elsewhere.py:5 module global ``_integrator``,              invariance.integrate``
Everywhere.py:326)
``xpsi.Signal.energy_integrator`` (Signal.py:10)            ``xpsi_b200.tools.energy_integrator``
``xpsi.Signal.phase_integrator`` (Signal.py:11)             ``xpsi_b200.tools.phase_integrator``
``xpsi.Instrument.Instrument.__call__`` (Instrument.py:     ``xpsi_b200.instrument.fold`` on
146-197)                                                    ``self.construct_matrix()``
``xpsi.likelihoods.default_background_marginalisation``     ``xpsi_b200.likelihoods.eval_marginal_likelihood``
``.eval_marginal_likelihood / .precomputation`` and the     ``/ precomputation``
same names inside user modules that imported them
(examples_fast/Modules/CustomSignal.py:7-8)
``xpsi.tools.{phase_interpolator, energy_interpolator,      the ``xpsi_b200.tools`` callables of the same names
synthesise_*}``, ``xpsi.surface_radiation_field.intensity``
==========================================================  ==================================================

Objects constructed *before* the call keep the callables they already hold (``HotRegion._integrator`` is an
instance attribute): pass them as ``instances`` to have them rebound too.  ``install`` returns a handle whose
``uninstall()`` restores every name.  Nothing here imports the reference or the oracle: the package object is
handed in by the caller, and everything bound computes on the GPU (no CPU fallback; the import of this module
fails when ``libxpsi_b200.so`` is missing).
"""
import importlib
import sys

from . import instrument as _instrument
from . import likelihoods as _likelihoods
from . import surface_radiation_field as _srf
from . import tools as _tools
from .cellmesh import integrator as _general
from .cellmesh import integrator_for_azimuthal_invariance as _azinv
from .cellmesh import integrator_for_time_invariance as _tinv


class Installed:
    def __init__(self):
        self._undo = []

    def _set(self, obj, name, value):
        missing = object()
        old = getattr(obj, name, missing)
        self._undo.append((obj, name, old, missing))
        setattr(obj, name, value)

    def uninstall(self):
        for obj, name, old, missing in reversed(self._undo):
            if old is missing:
                delattr(obj, name)
            else:
                setattr(obj, name, old)
        self._undo = []


def _instrument_call(self, signal, irange, orange):
    """xpsi/Instrument.py:146-197 with the contraction on the GPU."""
    self._cached_signal = _instrument.fold(self.construct_matrix(), signal, irange, orange)
    return self._cached_signal


def install(xpsi, user_modules=(), instances=()):
    """Rebind the hot-path seams of the imported reference package ``xpsi``.

    :param user_modules: modules that did ``from xpsi.likelihoods... import eval_marginal_likelihood`` (or any
        other seam name) at import time, e.g. a ``CustomSignal`` module; the names they hold are rebound.
    :param instances: already constructed ``HotRegion`` / ``Everywhere`` / ``Elsewhere`` objects.
    """
    h = Installed()
    pkg = xpsi.__name__

    def mod(name):
        # late-bound seams (HotRegion.py:551-572 imports its integrator inside a setter) may not be loaded yet
        try:
            return importlib.import_module(pkg + "." + name)
        except ImportError:
            return None

    seams = {
        "cellmesh.integrator_for_azimuthal_invariance": {"integrate": _azinv.integrate},
        "cellmesh.integrator": {"integrate": _general.integrate},
        "cellmesh.integrator_for_time_invariance": {"integrate": _tinv.integrate},
        "tools.energy_integrator": {"energy_integrator": _tools.energy_integrator},
        "tools.phase_integrator": {"phase_integrator": _tools.phase_integrator},
        "tools.phase_interpolator": {"phase_interpolator": _tools.phase_interpolator},
        "tools.energy_interpolator": {"energy_interpolator": _tools.energy_interpolator},
        "tools.synthesise": {"synthesise_exposure": _tools.synthesise_exposure,
                             "synthesise_given_total_count_number": _tools.synthesise_given_total_count_number},
        "tools": {"energy_integrator": _tools.energy_integrator, "phase_integrator": _tools.phase_integrator,
                  "phase_interpolator": _tools.phase_interpolator,
                  "energy_interpolator": _tools.energy_interpolator,
                  "synthesise_exposure": _tools.synthesise_exposure,
                  "synthesise_given_total_count_number": _tools.synthesise_given_total_count_number},
        "likelihoods.default_background_marginalisation": {
            "eval_marginal_likelihood": _likelihoods.eval_marginal_likelihood,
            "precomputation": _likelihoods.precomputation},
        "likelihoods._poisson_likelihood_given_background": {
            "poisson_likelihood_given_background": _likelihoods.poisson_likelihood_given_background},
        "likelihoods": {"eval_marginal_likelihood": _likelihoods.eval_marginal_likelihood,
                        "precomputation": _likelihoods.precomputation,
                        "poisson_likelihood_given_background": _likelihoods.poisson_likelihood_given_background},
        # names imported at module top by the classes that call them
        "Elsewhere": {"_integrator": _tinv.integrate},
        "Signal": {"energy_integrator": _tools.energy_integrator, "phase_integrator": _tools.phase_integrator},
        "surface_radiation_field": {"intensity": _srf.intensity},
    }
    originals = {}
    for name, table in seams.items():
        m = mod(name)
        if m is None:
            continue
        for attr, fn in table.items():
            if hasattr(m, attr):
                originals[id(getattr(m, attr))] = fn
                h._set(m, attr, fn)
    inst_mod = mod("Instrument")
    if inst_mod is not None:
        h._set(inst_mod.Instrument, "__call__", _instrument_call)
    # user modules: whatever reference callable they captured by ``from ... import`` is swapped for its mirror
    for um in user_modules:
        for attr, val in list(vars(um).items()):
            fn = originals.get(id(val))
            if fn is not None:
                h._set(um, attr, fn)
    for obj in instances:
        cur = getattr(obj, "_integrator", None)
        fn = originals.get(id(cur))
        if fn is not None:
            h._set(obj, "_integrator", fn)
    # the global interpolant switches stay in step with the reference's (xpsi/tools/core.pyx:34-82)
    try:
        ref_tools = sys.modules[pkg + ".tools"]
        _tools.set_phase_interpolant(ref_tools.get_phase_interpolant())
        _tools.set_energy_interpolant(ref_tools.get_energy_interpolant())
    except Exception:
        pass
    return h
