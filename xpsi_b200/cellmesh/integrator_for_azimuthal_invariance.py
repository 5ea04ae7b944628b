"""GPU drop-in for ``xpsi.cellmesh.integrator_for_azimuthal_invariance``."""
import numpy as np

from .. import _lib
from ..tools import phase_interpolant_id


def integrate(numThreads, R, omega, r_s, inclination, cellArea, radialCoords_of_parallels,
              r_s_over_r, theta, phi, srcCellParams, CELL_RADIATES, correction_srcCellParams,
              numRays, deflection, cos_alpha, lag, maxDeflection, cos_gammaArray, energies, leaves,
              phases, hot_atmosphere, elsewhere_atmosphere, hot_atm_ext, else_atm_ext, beam_opt,
              image_order_limit=None, R_in=1e6):
    """Same positional signature and return convention as
    xpsi/cellmesh/integrator_for_azimuthal_invariance.pyx:70-98:
    ``(0, flux[N_E, N_P])`` on success, ``(1, None)`` on a numerical error.
    ``numThreads`` is accepted and ignored.
    """
    return _integrate(_lib.lib.xpsi_b200_integrate_azimuthal_invariance, R, omega, r_s, inclination, cellArea,
                      radialCoords_of_parallels, r_s_over_r, theta, phi, srcCellParams, CELL_RADIATES,
                      correction_srcCellParams, numRays, deflection, cos_alpha, lag, maxDeflection,
                      cos_gammaArray, energies, leaves, phases, hot_atmosphere, elsewhere_atmosphere,
                      hot_atm_ext, else_atm_ext, beam_opt, image_order_limit, R_in)


def _integrate(entry, R, omega, r_s, inclination, cellArea, radialCoords_of_parallels, r_s_over_r, theta, phi,
               srcCellParams, CELL_RADIATES, correction_srcCellParams, numRays, deflection, cos_alpha, lag,
               maxDeflection, cos_gammaArray, energies, leaves, phases, hot_atmosphere, elsewhere_atmosphere,
               hot_atm_ext, else_atm_ext, beam_opt, image_order_limit, R_in):
    """Marshal one hot-region member to a C-ABI pulse integrator (shared with ``cellmesh.integrator``)."""
    cellArea = _lib.as_f8(cellArea, 2)
    n_rings, n_azi = cellArea.shape
    theta = _lib.as_f8(theta, 2)
    phi = _lib.as_f8(phi, 2)
    srcCellParams = _lib.as_f8(srcCellParams, 3)
    CELL_RADIATES = _lib.as_i4(CELL_RADIATES, 2)
    radial = _lib.as_f8(radialCoords_of_parallels, 1)
    rsr = _lib.as_f8(r_s_over_r, 1)
    deflection = _lib.as_f8(deflection, 2)
    cos_alpha = _lib.as_f8(cos_alpha, 2)
    lag = _lib.as_f8(lag, 2)
    maxDeflection = _lib.as_f8(maxDeflection, 1)
    cos_gamma = _lib.as_f8(cos_gammaArray, 1)
    energies = _lib.as_f8(energies, 1)
    leaves = _lib.as_f8(leaves, 1)
    phases = _lib.as_f8(phases, 1)
    corr = None
    if correction_srcCellParams is not None:
        corr = _lib.as_f8(correction_srcCellParams, 3)
    hot = _lib.Atmosphere.get(hot_atmosphere)
    els = _lib.Atmosphere.get(elsewhere_atmosphere) if corr is not None else None
    flux = np.zeros((energies.shape[0], phases.shape[0]), dtype=np.float64)
    rc = entry(
        float(R), float(omega), float(r_s), float(inclination), n_rings, n_azi,
        _lib.dptr(cellArea), _lib.dptr(radial), _lib.dptr(rsr), _lib.dptr(theta), _lib.dptr(phi),
        _lib.dptr(srcCellParams), srcCellParams.shape[2], _lib.iptr(CELL_RADIATES),
        _lib.dptr(corr) if corr is not None else None,
        int(numRays), _lib.dptr(deflection), _lib.dptr(cos_alpha), _lib.dptr(lag),
        _lib.dptr(maxDeflection), _lib.dptr(cos_gamma),
        energies.shape[0], _lib.dptr(energies), leaves.shape[0], _lib.dptr(leaves),
        phases.shape[0], _lib.dptr(phases),
        hot.handle if hot is not None else None, els.handle if els is not None else None,
        int(hot_atm_ext), int(else_atm_ext) if else_atm_ext is not None else 0, int(beam_opt),
        int(image_order_limit) if image_order_limit is not None else 0, float(R_in),
        phase_interpolant_id(), _lib.dptr(flux))
    if rc == _lib.ENUMERICAL:
        return (1, None)
    if rc == _lib.EUNSUPPORTED:
        raise NotImplementedError("xpsi_b200: " + _lib.last_error())
    _lib.check(rc)
    return (0, flux)
