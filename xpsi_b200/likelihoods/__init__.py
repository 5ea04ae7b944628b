"""GPU drop-ins for ``xpsi.likelihoods.default_background_marginalisation``."""
import ctypes as C

import numpy as np

from .. import _lib
from ..tools import phase_interpolant_id


def precomputation(data):
    """-sum_j ln(d_ij!) per channel
    (xpsi/likelihoods/default_background_marginalisation.pyx:38-68)."""
    data = _lib.as_i4(data, 2)
    out = np.empty(data.shape[0], dtype=np.float64)
    _lib.check(_lib.lib.xpsi_b200_precomputation(_lib.iptr(data), data.shape[0], data.shape[1],
                                                 _lib.dptr(out)))
    return out


def eval_marginal_likelihood(exposure_time, phases, counts, components, component_phases,
                             phase_shifts, neg_sum_ln_data_factorial, support,
                             workspace_intervals, epsabs, epsrel, epsilon, sigmas, llzero,
                             allow_negative=False, slim=20.0, background=None):
    """Same signature and 4-tuple return as
    xpsi/likelihoods/default_background_marginalisation.pyx:450-466.

    ``workspace_intervals``, ``epsabs`` and ``epsrel`` configure GSL CQUAD in the
    reference; the GPU quadrature is fixed-order and converged below 1e-12, so
    they are accepted and ignored.  On the two paths where the reference
    returns a *random* near-``llzero`` number (slim early exit, non-positive
    integral) the same convention is followed.
    """
    phases = _lib.as_f8(phases, 1)
    counts = _lib.as_f8(counts, 2)
    support = _lib.as_f8(support, 2)
    precomp = _lib.as_f8(neg_sum_ln_data_factorial, 1)
    comps = [_lib.as_f8(c, 2) for c in components]
    cph = [_lib.as_f8(p, 1) for p in component_phases]
    shifts = _lib.as_f8(phase_shifts, 1)
    n_chan = counts.shape[0]
    if support.shape[0] != n_chan:
        raise TypeError('The number of energy channels in the background support does not match to that of the data.')
    if (support[:, 1] == 0).any():
        raise TypeError('Background upper limit cannot be set to 0.')
    if ((support[:, 1] > 0) & (support[:, 1] - support[:, 0] < 0)).any():
        raise TypeError('Background upper limit must be higher than the lower limit.')
    if counts.shape[1] != phases.shape[0] - 1 or precomp.shape[0] != n_chan or shifts.shape[0] != len(comps):
        raise ValueError("counts / phases / neg_sum_ln_data_factorial / phase_shifts shapes do not match")
    arr, parr, nph = _component_arrays(comps, cph, n_chan)
    allow = _allow_flags(allow_negative, len(comps))
    bg = _lib.as_f8(background, 2) if background is not None else None
    if bg is not None and bg.shape != counts.shape:
        raise ValueError("background must have the shape of the data")
    n_bins = phases.shape[0] - 1
    lnL = C.c_double(0.0)
    star = np.zeros((n_chan, n_bins), dtype=np.float64)
    mcl = np.zeros(n_chan, dtype=np.float64)
    mcl_s = np.zeros(n_chan, dtype=np.float64)
    rc = _lib.lib.xpsi_b200_eval_marginal_likelihood(
        float(exposure_time), _lib.dptr(phases), n_bins, _lib.dptr(counts), n_chan, arr, len(comps),
        parr, _lib.iptr(nph), _lib.dptr(shifts), _lib.dptr(precomp), _lib.dptr(support),
        float(epsilon), float(sigmas), float(llzero), _lib.iptr(allow), float(slim),
        _lib.dptr(bg) if bg is not None else None, phase_interpolant_id(),
        C.cast(C.pointer(lnL), _lib.c_double_p), _lib.dptr(star), _lib.dptr(mcl), _lib.dptr(mcl_s))
    if rc in (_lib.ESLIM, _lib.EQUADRATURE):
        return (llzero * (0.1 + 0.9 * np.random.rand()), star, mcl, mcl_s)
    if rc == _lib.EUNSUPPORTED:
        raise NotImplementedError("xpsi_b200: " + _lib.last_error())
    _lib.check(rc)
    return (lnL.value, star, mcl, mcl_s)


def _component_arrays(comps, cph, n_chan):
    """Pointer tables of the components and of their phase grids (one grid per component, as
    compute_expected_counts.pyx:66-197 takes them) and the node counts."""
    if len(comps) != len(cph):
        raise ValueError("one phase grid per component is required")
    for c, p in zip(comps, cph):
        if c.shape != (n_chan, p.shape[0]):
            raise ValueError("a component of shape %r does not match %d channels x %d phases"
                             % (c.shape, n_chan, p.shape[0]))
    arr = (_lib.c_double_p * len(comps))(*[_lib.dptr(c) for c in comps])
    parr = (_lib.c_double_p * len(cph))(*[_lib.dptr(p) for p in cph])
    nph = np.array([p.shape[0] for p in cph], dtype=np.int32)
    return arr, parr, nph


def _allow_flags(allow_negative, n_comp):
    """One flag per component (default_background_marginalisation.pyx:497-508)."""
    if isinstance(allow_negative, (bool, np.bool_, int)):
        return np.full(n_comp, int(bool(allow_negative)), dtype=np.int32)
    vals = [int(bool(v)) for v in allow_negative]
    if len(vals) != n_comp:
        raise ValueError('Number of allow_negative declarations does not match the number of components..')
    return np.array(vals, dtype=np.int32)


def _components(components, component_phases, n_chan):
    comps = [_lib.as_f8(c, 2) for c in components]
    cph = [_lib.as_f8(p, 1) for p in component_phases]
    return comps, cph


def poisson_likelihood_given_background(exposure_time, phases, counts, components, component_phases,
                                        phase_shifts, background, neg_sum_ln_data_factorial=None,
                                        allow_negative=False):
    """Poisson likelihood for a given background (count rate), same signature and ``(lnL, expected)``
    return as xpsi/likelihoods/_poisson_likelihood_given_background.pyx:14-113."""
    phases = _lib.as_f8(phases, 1)
    counts = _lib.as_f8(counts, 2)
    bg = _lib.as_f8(background, 2)
    comps, cph = _components(components, component_phases, counts.shape[0])
    shifts = _lib.as_f8(phase_shifts, 1)
    pre = _lib.as_f8(neg_sum_ln_data_factorial, 1) if neg_sum_ln_data_factorial is not None else None
    n_chan, n_bins = counts.shape
    if bg.shape != counts.shape or phases.shape[0] != n_bins + 1 or shifts.shape[0] != len(comps) or \
            (pre is not None and pre.shape[0] != n_chan):
        raise ValueError("counts / background / phases / phase_shifts shapes do not match")
    arr, parr, nph = _component_arrays(comps, cph, n_chan)
    allow = _allow_flags(allow_negative, len(comps))
    lnL = C.c_double(0.0)
    expec = np.zeros((n_chan, n_bins), dtype=np.float64)
    rc = _lib.lib.xpsi_b200_poisson_likelihood_given_background(
        float(exposure_time), _lib.dptr(phases), n_bins, _lib.dptr(counts), n_chan, arr, len(comps),
        parr, _lib.iptr(nph), _lib.dptr(shifts), _lib.dptr(bg), _lib.dptr(pre) if pre is not None else None,
        _lib.iptr(allow), phase_interpolant_id(), C.cast(C.pointer(lnL), _lib.c_double_p), _lib.dptr(expec))
    if rc == _lib.EQUADRATURE:
        return (-1.0e90 * (0.1 + 0.9 * np.random.rand()), expec)
    if rc == _lib.EUNSUPPORTED:
        raise NotImplementedError("xpsi_b200: " + _lib.last_error())
    _lib.check(rc)
    return (lnL.value, expec)


def expected_counts(exposure_time, phases, components, component_phases, phase_shifts, background,
                    allow_negative=False):
    """Expected counts ``T (star + background)`` per (channel, phase bin): the deterministic half of
    ``tools.synthesise_exposure`` (xpsi/tools/compute_expected_counts.pyx:200-315)."""
    phases = _lib.as_f8(phases, 1)
    bg = _lib.as_f8(background, 2)
    comps, cph = _components(components, component_phases, bg.shape[0])
    shifts = _lib.as_f8(phase_shifts, 1)
    if bg.shape[1] != phases.shape[0] - 1 or shifts.shape[0] != len(comps):
        raise ValueError("background / phases / phase_shifts shapes do not match")
    arr, parr, nph = _component_arrays(comps, cph, bg.shape[0])
    allow = _allow_flags(allow_negative, len(comps))
    expec = np.zeros(bg.shape, dtype=np.float64)
    rc = _lib.lib.xpsi_b200_poisson_likelihood_given_background(
        float(exposure_time), _lib.dptr(phases), phases.shape[0] - 1, None, bg.shape[0], arr, len(comps),
        parr, _lib.iptr(nph), _lib.dptr(shifts), _lib.dptr(bg), None, _lib.iptr(allow),
        phase_interpolant_id(), None, _lib.dptr(expec))
    _lib.check(rc)
    return expec
