"""GPU drop-in for ``xpsi.tools.synthesise`` (expected counts on the device, Poisson draw on the host)."""
import numpy as np

from ..likelihoods import expected_counts


def _poisson(expected, gsl_seed):
    # The reference draws with GSL's default generator (xpsi/tools/synthesise.pyx:122-141); GSL is not a
    # dependency here, so the realisation comes from numpy's generator with the same seed argument: the
    # expected counts are the reference's, the noise stream is not.
    rng = np.random.default_rng(gsl_seed)
    return rng.poisson(np.clip(expected, 0.0, None)).astype(np.float64)


def synthesise_exposure(exposure_time, phases, components, component_phases, phase_shifts,
                        expected_background_counts, background, allow_negative=False, gsl_seed=None):
    """Same signature and ``(expected, synthetic, background scale)`` return as
    xpsi/tools/synthesise.pyx:40-149."""
    phases = np.ascontiguousarray(phases, dtype=np.float64)
    background = np.ascontiguousarray(background, dtype=np.float64)
    n_chan, n_bins = components[0].shape[0], phases.shape[0] - 1
    BACKGROUND = float(np.sum(background[:n_chan, :n_bins]))
    SCALE_BACKGROUND = 0.0 if BACKGROUND == 0.0 else expected_background_counts / BACKGROUND
    rate = np.ascontiguousarray(background[:n_chan, :n_bins] * SCALE_BACKGROUND / exposure_time)
    EXPEC = expected_counts(exposure_time, phases, components, component_phases, phase_shifts, rate,
                            allow_negative)
    return (EXPEC, _poisson(EXPEC, gsl_seed), SCALE_BACKGROUND)


def synthesise_given_total_count_number(phases, expected_star_counts, components, component_phases, phase_shifts,
                                        expected_background_counts, background, allow_negative=False,
                                        gsl_seed=None):
    """Same signature and 4-tuple return as xpsi/tools/synthesise.pyx:152-276: ``(expected, synthetic,
    star scale, background scale / star scale)``."""
    phases = np.ascontiguousarray(phases, dtype=np.float64)
    background = np.ascontiguousarray(background, dtype=np.float64)
    n_chan, n_bins = components[0].shape[0], phases.shape[0] - 1
    STAR = expected_counts(1e5, phases, components, component_phases, phase_shifts,
                           np.zeros((n_chan, n_bins), dtype=np.float64), allow_negative)
    STAR_total = float(np.sum(STAR))
    BACKGROUND = float(np.sum(background[:n_chan, :n_bins]))
    SCALE_STAR = expected_star_counts / STAR_total
    SCALE_BACKGROUND = 0.0 if BACKGROUND == 0.0 else expected_background_counts / BACKGROUND
    EXPEC = np.ascontiguousarray(STAR * SCALE_STAR + background[:n_chan, :n_bins] * SCALE_BACKGROUND)
    return (EXPEC, _poisson(EXPEC, gsl_seed), SCALE_STAR, SCALE_BACKGROUND / SCALE_STAR)
