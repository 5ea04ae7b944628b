"""Mirror of ``xpsi.tools`` for the hot path (xpsi/tools/core.pyx, energy_integrator.pyx)."""
import numpy as np

from .. import _lib

__interpolants__ = _lib.INTERPOLANTS
_phase_interpolant = 0      # default Akima periodic, xpsi/__init__.py:55
_energy_interpolant = 1     # default Steffen,        xpsi/__init__.py:56


def set_phase_interpolant(interpolant):
    """xpsi/tools/core.pyx:34-52."""
    global _phase_interpolant
    if not isinstance(interpolant, str):
        raise TypeError('Interpolant declaration must be a string.')
    if interpolant not in __interpolants__:
        raise ValueError('Invalid interpolant name. See the docstring.')
    _phase_interpolant = __interpolants__[interpolant]


def get_phase_interpolant():
    for key, value in __interpolants__.items():
        if value == _phase_interpolant:
            return key


def set_energy_interpolant(interpolant):
    """xpsi/tools/core.pyx:64-82."""
    global _energy_interpolant
    if not isinstance(interpolant, str):
        raise TypeError('Interpolant declaration must be a string.')
    if interpolant not in __interpolants__:
        raise ValueError('Invalid interpolant name. See the docstring.')
    _energy_interpolant = __interpolants__[interpolant]


def get_energy_interpolant():
    for key, value in __interpolants__.items():
        if value == _energy_interpolant:
            return key


def phase_interpolant_id():
    return _phase_interpolant


def energy_integrator(N_Ts, signal, energies, energy_edges):
    """Integrate a signal over energy intervals on the GPU.

    Same signature and return layout as xpsi/tools/energy_integrator.pyx:27-114:
    ``signal[N_E, N_P]``, ``energies`` and ``energy_edges`` are log10 values, the
    result is ``[len(energy_edges)-1, N_P]``.  ``N_Ts`` (OpenMP threads) is
    accepted and ignored.
    """
    signal = _lib.as_f8(signal, 2)
    energies = _lib.as_f8(energies, 1)
    energy_edges = _lib.as_f8(energy_edges, 1)
    if signal.shape[0] != energies.shape[0]:
        raise ValueError("signal rows must match the number of energies")
    n_in = energy_edges.shape[0] - 1
    out = np.empty((n_in, signal.shape[1]), dtype=np.float64)
    _lib.check(_lib.lib.xpsi_b200_energy_integrator(
        _lib.dptr(signal), signal.shape[0], signal.shape[1], _lib.dptr(energies),
        _lib.dptr(energy_edges), n_in, _phase_interpolant, _lib.dptr(out)))
    return out


def energy_interpolator(N_Ts, signal, energies, new_energies):
    """Interpolate a signal in energy (xpsi/tools/energy_interpolator.pyx:27-125): spline in log10 E
    with the global *energy* interpolant, in log10 of the signal when a column is strictly positive,
    zero above the last energy.  Returns ``[len(new_energies), N_P]``."""
    signal = _lib.as_f8(signal, 2)
    energies = _lib.as_f8(energies, 1)
    new_energies = _lib.as_f8(new_energies, 1)
    out = np.empty((new_energies.shape[0], signal.shape[1]), dtype=np.float64)
    _lib.check(_lib.lib.xpsi_b200_energy_interpolator(
        _lib.dptr(signal), signal.shape[0], signal.shape[1], _lib.dptr(energies), _lib.dptr(new_energies),
        new_energies.shape[0], _energy_interpolant, _lib.dptr(out)))
    return out


def phase_integrator(exposure_time, phases, signal, signal_phases, phase_shift, allow_negative=0):
    """Integrate a phase-shifted signal over phase intervals and scale by the exposure time
    (xpsi/tools/phase_integrator.pyx:23-121).  Returns ``[rows, len(phases)-1]``."""
    phases = _lib.as_f8(phases, 1)
    signal = _lib.as_f8(signal, 2)
    signal_phases = _lib.as_f8(signal_phases, 1)
    out = np.empty((signal.shape[0], phases.shape[0] - 1), dtype=np.float64)
    _lib.check(_lib.lib.xpsi_b200_phase_integrator(
        float(exposure_time), _lib.dptr(phases), phases.shape[0] - 1, _lib.dptr(signal), signal.shape[0],
        _lib.dptr(signal_phases), signal_phases.shape[0], float(phase_shift), int(bool(allow_negative)),
        _phase_interpolant, _lib.dptr(out)))
    return out


def phase_interpolator(new_phases, phases, signal, phase_shift, allow_negative=0):
    """Interpolate a phase-shifted signal at new phases (xpsi/tools/phase_interpolator.pyx:25-98).
    Returns ``[rows, len(new_phases)]``."""
    new_phases = _lib.as_f8(new_phases, 1)
    phases = _lib.as_f8(phases, 1)
    signal = _lib.as_f8(signal, 2)
    out = np.empty((signal.shape[0], new_phases.shape[0]), dtype=np.float64)
    _lib.check(_lib.lib.xpsi_b200_phase_interpolator(
        _lib.dptr(new_phases), new_phases.shape[0], _lib.dptr(phases), phases.shape[0], _lib.dptr(signal),
        signal.shape[0], float(phase_shift), int(bool(allow_negative)), _phase_interpolant, _lib.dptr(out)))
    return out


def synthesise_exposure(*args, **kwargs):
    """xpsi.tools.synthesise_exposure (xpsi/tools/synthesise.pyx:40-149); see tools/synthesise.py."""
    from .synthesise import synthesise_exposure as f
    return f(*args, **kwargs)


def synthesise_given_total_count_number(*args, **kwargs):
    """xpsi.tools.synthesise_given_total_count_number (xpsi/tools/synthesise.pyx:152-276)."""
    from .synthesise import synthesise_given_total_count_number as f
    return f(*args, **kwargs)
