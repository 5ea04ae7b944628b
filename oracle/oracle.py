"""ctypes front-end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by xpsi_b200/.
Function signatures mirror the reference's Python callables so the parity tests
read like the reference's own.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def _load():
    if not os.path.isfile(_LIB):
        subprocess.check_call(["make", "-C", _HERE])
    return C.CDLL(_LIB)


lib = _load()


def _d(a):
    return a.ctypes.data_as(dp)


def _f8(a):
    return np.ascontiguousarray(a, dtype=np.float64)


INTERP = {'Akima': 0, 'Steffen': 1, 'Cubic': 2}


def integrate(numThreads, R, omega, r_s, inclination, cellArea, radialCoords_of_parallels, r_s_over_r,
              theta, phi, srcCellParams, CELL_RADIATES, correction_srcCellParams, numRays, deflection,
              cos_alpha, lag, maxDeflection, cos_gammaArray, energies, leaves, phases, hot_atmosphere,
              elsewhere_atmosphere, hot_atm_ext, else_atm_ext, beam_opt, image_order_limit=None,
              R_in=1e6, phase_interpolant='Akima'):
    """xpsi/cellmesh/integrator_for_azimuthal_invariance.pyx:70-98 (no disc, beam_opt 0)."""
    assert beam_opt in (0, 1, 2, 3)
    cellArea, theta, phi = _f8(cellArea), _f8(theta), _f8(phi)
    par = _f8(srcCellParams)
    rad = np.ascontiguousarray(CELL_RADIATES, dtype=np.int32)
    arrs = [_f8(x) for x in (radialCoords_of_parallels, r_s_over_r, deflection, cos_alpha, lag,
                             maxDeflection, cos_gammaArray, energies, leaves, phases)]
    radial, rsr, defl, ca, lg, maxd, cg, E, L, P = arrs
    if hot_atmosphere:
        tab = [_f8(t) for t in hot_atmosphere]
    else:
        tab = [np.zeros(4)] * 5
    flux = np.zeros((E.size, P.size))
    rc = lib.oracle_integrate_azinv(
        C.c_double(omega), C.c_double(inclination), C.c_int(cellArea.shape[0]), C.c_int(cellArea.shape[1]),
        _d(cellArea), _d(radial), _d(rsr), _d(theta), _d(phi), _d(par), C.c_int(par.shape[2]),
        rad.ctypes.data_as(ip), C.c_int(int(numRays)), _d(defl), _d(ca), _d(lg), _d(maxd), _d(cg),
        C.c_int(E.size), _d(E), C.c_int(L.size), _d(L), C.c_int(P.size), _d(P), C.c_int(int(hot_atm_ext)),
        _d(tab[0]), C.c_int(tab[0].size), _d(tab[1]), C.c_int(tab[1].size), _d(tab[2]), C.c_int(tab[2].size),
        _d(tab[3]), C.c_int(tab[3].size), _d(tab[4]),
        C.c_int(int(image_order_limit) if image_order_limit else 0), C.c_int(INTERP[phase_interpolant]), _d(flux),
        *(_correction_args(correction_srcCellParams, elsewhere_atmosphere, else_atm_ext) +
          [C.c_double(float(R_in)), C.c_int(int(beam_opt))]))
    return (1, None) if rc else (0, flux)


def integrate_general(numThreads, R, omega, r_s, inclination, cellArea, radialCoords_of_parallels, r_s_over_r,
                      theta, phi, srcCellParams, CELL_RADIATES, correction_srcCellParams, numRays, deflection,
                      cos_alpha, lag, maxDeflection, cos_gammaArray, energies, leaves, phases, hot_atmosphere,
                      elsewhere_atmosphere, hot_atm_ext, else_atm_ext, beam_opt, image_order_limit=None,
                      R_in=1e6, phase_interpolant='Akima'):
    """xpsi/cellmesh/integrator.pyx:48-76 (the integrator without azimuthal invariance)."""
    assert beam_opt in (0, 1, 2, 3)
    cellArea, theta, phi = _f8(cellArea), _f8(theta), _f8(phi)
    par = _f8(srcCellParams)
    rad = np.ascontiguousarray(CELL_RADIATES, dtype=np.int32)
    arrs = [_f8(x) for x in (radialCoords_of_parallels, r_s_over_r, deflection, cos_alpha, lag,
                             maxDeflection, cos_gammaArray, energies, leaves, phases)]
    radial, rsr, defl, ca, lg, maxd, cg, E, L, P = arrs
    tab = [_f8(t) for t in hot_atmosphere] if hot_atmosphere else [np.zeros(4)] * 5
    flux = np.zeros((E.size, P.size))
    rc = lib.oracle_integrate_general(
        C.c_double(omega), C.c_double(inclination), C.c_int(cellArea.shape[0]), C.c_int(cellArea.shape[1]),
        _d(cellArea), _d(radial), _d(rsr), _d(theta), _d(phi), _d(par), C.c_int(par.shape[2]),
        rad.ctypes.data_as(ip), C.c_int(int(numRays)), _d(defl), _d(ca), _d(lg), _d(maxd), _d(cg),
        C.c_int(E.size), _d(E), C.c_int(L.size), _d(L), C.c_int(P.size), _d(P), C.c_int(int(hot_atm_ext)),
        _d(tab[0]), C.c_int(tab[0].size), _d(tab[1]), C.c_int(tab[1].size), _d(tab[2]), C.c_int(tab[2].size),
        _d(tab[3]), C.c_int(tab[3].size), _d(tab[4]),
        C.c_int(int(image_order_limit) if image_order_limit else 0), C.c_int(INTERP[phase_interpolant]), _d(flux),
        *(_correction_args(correction_srcCellParams, elsewhere_atmosphere, else_atm_ext) +
          [C.c_int(int(beam_opt))]))
    return (1, None) if rc else (0, flux)


def intensity(energies, mu, local_variables, atmosphere=None, stokesQ=0, region_extension='hot',
              atmos_extension="BB", beam_opt=0, numTHREADS=1):
    """xpsi/surface_radiation_field/core.pyx:125-308 ('BB' and 'Num4D', Stokes I)."""
    assert stokesQ == 0 and atmos_extension in ("BB", "Num4D")
    E, m, v = _f8(energies), _f8(mu), _f8(local_variables)
    tab = [_f8(t) for t in atmosphere] if atmosphere else [np.zeros(4)] * 5
    out = np.zeros(E.size)
    lib.oracle_intensity(C.c_int(E.size), _d(E), _d(m), _d(v), C.c_int(v.shape[1]),
                         C.c_int(0 if region_extension == 'hot' else 1), C.c_int(1 if atmos_extension == "BB" else 2),
                         _d(tab[0]), C.c_int(tab[0].size), _d(tab[1]), C.c_int(tab[1].size), _d(tab[2]),
                         C.c_int(tab[2].size), _d(tab[3]), C.c_int(tab[3].size), _d(tab[4]), C.c_int(int(beam_opt)),
                         _d(out))
    return out


def _correction_args(correction, atmosphere, else_atm_ext):
    if correction is None:
        z = np.zeros(4)
        return [None, C.c_int(0)] + [_d(z), C.c_int(4)] * 4 + [_d(z)]
    corr = _f8(correction)
    tab = [_f8(t) for t in atmosphere] if atmosphere else [np.zeros(4)] * 5
    _correction_args.keep = (corr, tab)
    return [_d(corr), C.c_int(int(else_atm_ext)), _d(tab[0]), C.c_int(tab[0].size), _d(tab[1]), C.c_int(tab[1].size),
            _d(tab[2]), C.c_int(tab[2].size), _d(tab[3]), C.c_int(tab[3].size), _d(tab[4])]


def energy_integrator(N_Ts, signal, energies, energy_edges, phase_interpolant='Akima'):
    """xpsi/tools/energy_integrator.pyx:27-114."""
    signal, energies, energy_edges = _f8(signal), _f8(energies), _f8(energy_edges)
    out = np.zeros((energy_edges.size - 1, signal.shape[1]))
    lib.oracle_energy_integrator(_d(signal), C.c_int(signal.shape[0]), C.c_int(signal.shape[1]), _d(energies),
                                 _d(energy_edges), C.c_int(energy_edges.size - 1),
                                 C.c_int(INTERP[phase_interpolant]), _d(out))
    return out


def fold(matrix, signal):
    """numpy.dot(matrix, signal), xpsi/Instrument.py:192-197."""
    matrix, signal = _f8(matrix), _f8(signal)
    out = np.zeros((matrix.shape[0], signal.shape[1]))
    lib.oracle_fold(_d(matrix), C.c_int(matrix.shape[0]), C.c_int(matrix.shape[1]), _d(signal),
                    C.c_int(signal.shape[1]), _d(out))
    return out


def precomputation(data):
    data = np.ascontiguousarray(data, dtype=np.int32)
    out = np.zeros(data.shape[0])
    lib.oracle_precomputation(data.ctypes.data_as(ip), C.c_int(data.shape[0]), C.c_int(data.shape[1]), _d(out))
    return out


def eval_marginal_likelihood(exposure_time, phases, counts, components, component_phases, phase_shifts,
                             neg_sum_ln_data_factorial, support, workspace_intervals, epsabs, epsrel,
                             epsilon, sigmas, llzero, allow_negative=False, slim=20.0, background=None,
                             phase_interpolant='Akima'):
    """xpsi/likelihoods/default_background_marginalisation.pyx:450-734.
    Returns (status, lnL, expected counts, ML background, ML background given support)."""
    phases, counts, support = _f8(phases), _f8(counts), _f8(support)
    comps = [_f8(c) for c in components]
    cph = _f8(component_phases[0])
    shifts, pre = _f8(phase_shifts), _f8(neg_sum_ln_data_factorial)
    bg = _f8(background) if background is not None else None
    arr = (dp * len(comps))(*[_d(c) for c in comps])
    lnL = C.c_double(0.0)
    n_chan, n_bins = counts.shape
    star, mcl, mcls = np.zeros((n_chan, n_bins)), np.zeros(n_chan), np.zeros(n_chan)
    rc = lib.oracle_eval_marginal_likelihood(
        C.c_double(exposure_time), _d(phases), C.c_int(n_bins), _d(counts), C.c_int(n_chan), arr,
        C.c_int(len(comps)), _d(cph), C.c_int(cph.size), _d(shifts), _d(pre), _d(support),
        C.c_int(int(workspace_intervals)), C.c_double(epsabs), C.c_double(epsrel), C.c_double(epsilon),
        C.c_double(sigmas), C.c_double(llzero), C.c_int(int(bool(allow_negative))), C.c_double(slim),
        _d(bg) if bg is not None else None, C.c_int(INTERP[phase_interpolant]), C.byref(lnL), _d(star),
        _d(mcl), _d(mcls))
    return rc, lnL.value, star, mcl, mcls


def phase_integrator(exposure_time, phases, signal, signal_phases, phase_shift, allow_negative=0,
                     phase_interpolant='Akima'):
    """xpsi/tools/phase_integrator.pyx:23-121."""
    phases, signal, signal_phases = _f8(phases), _f8(signal), _f8(signal_phases)
    out = np.zeros((signal.shape[0], phases.size - 1))
    lib.oracle_phase_integrator(C.c_double(exposure_time), _d(phases), C.c_int(phases.size - 1), _d(signal),
                                C.c_int(signal.shape[0]), _d(signal_phases), C.c_int(signal_phases.size),
                                C.c_double(phase_shift), C.c_int(int(bool(allow_negative))),
                                C.c_int(INTERP[phase_interpolant]), _d(out))
    return out


def phase_interpolator(new_phases, phases, signal, phase_shift, allow_negative=0, phase_interpolant='Akima'):
    """xpsi/tools/phase_interpolator.pyx:25-98."""
    new_phases, phases, signal = _f8(new_phases), _f8(phases), _f8(signal)
    out = np.zeros((signal.shape[0], new_phases.size))
    lib.oracle_phase_interpolator(_d(new_phases), C.c_int(new_phases.size), _d(phases), C.c_int(phases.size),
                                  _d(signal), C.c_int(signal.shape[0]), C.c_double(phase_shift),
                                  C.c_int(int(bool(allow_negative))), C.c_int(INTERP[phase_interpolant]), _d(out))
    return out


def energy_interpolator(N_Ts, signal, energies, new_energies, energy_interpolant='Steffen'):
    """xpsi/tools/energy_interpolator.pyx:27-125."""
    signal, energies, new_energies = _f8(signal), _f8(energies), _f8(new_energies)
    out = np.zeros((new_energies.size, signal.shape[1]))
    lib.oracle_energy_interpolator(_d(signal), C.c_int(signal.shape[0]), C.c_int(signal.shape[1]), _d(energies),
                                   _d(new_energies), C.c_int(new_energies.size),
                                   C.c_int(INTERP[energy_interpolant]), _d(out))
    return out


def integrate_time_invariance(numThreads, R, omega, r_s, inclination, sqrt_numPix, cellArea,
                              radialCoords_of_parallels, r_s_over_r, theta, phi, srcCellParams, numRays,
                              deflection, cos_alpha, maxDeflection, cos_gammaArray, energies, atmosphere,
                              atm_ext, image_order_limit=None, *args):
    """xpsi/cellmesh/integrator_for_time_invariance.pyx:59-80."""
    theta, phi, par = _f8(theta), _f8(phi), _f8(srcCellParams)
    radial, rsr, defl, ca, maxd, cg, E = [_f8(x) for x in (radialCoords_of_parallels, r_s_over_r, deflection,
                                                            cos_alpha, maxDeflection, cos_gammaArray, energies)]
    tab = [_f8(t) for t in atmosphere] if atmosphere else [np.zeros(4)] * 5
    flux = np.zeros(E.size)
    rc = lib.oracle_integrate_tinv(
        C.c_double(omega), C.c_double(inclination), C.c_int(int(sqrt_numPix)), C.c_double(cellArea), _d(radial),
        _d(rsr), _d(theta), _d(phi), _d(par), C.c_int(par.shape[2]), C.c_int(int(numRays)), _d(defl), _d(ca),
        _d(maxd), _d(cg), C.c_int(E.size), _d(E), C.c_int(int(atm_ext)), _d(tab[0]), C.c_int(tab[0].size),
        _d(tab[1]), C.c_int(tab[1].size), _d(tab[2]), C.c_int(tab[2].size), _d(tab[3]), C.c_int(tab[3].size),
        _d(tab[4]), C.c_int(int(image_order_limit) if image_order_limit else 0), _d(flux))
    return (1, None) if rc else (0, flux)
