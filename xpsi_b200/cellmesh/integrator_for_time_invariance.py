"""GPU drop-in for ``xpsi.cellmesh.integrator_for_time_invariance``."""
import numpy as np

from .. import _lib


def integrate(numThreads, R, omega, r_s, inclination, sqrt_numPix, cellArea, radialCoords_of_parallels,
              r_s_over_r, theta, phi, srcCellParams, numRays, deflection, cos_alpha, maxDeflection,
              cos_gammaArray, energies, atmosphere, atm_ext, image_order_limit=None, *args):
    """Same positional signature and return convention as
    xpsi/cellmesh/integrator_for_time_invariance.pyx:59-80: ``(0, flux[N_E])`` or ``(1, None)``.
    ``numThreads`` is accepted and ignored."""
    theta = _lib.as_f8(theta, 2)
    phi = _lib.as_f8(phi, 2)
    par = _lib.as_f8(srcCellParams, 3)
    radial = _lib.as_f8(radialCoords_of_parallels, 1)
    rsr = _lib.as_f8(r_s_over_r, 1)
    deflection = _lib.as_f8(deflection, 2)
    cos_alpha = _lib.as_f8(cos_alpha, 2)
    maxDeflection = _lib.as_f8(maxDeflection, 1)
    cos_gamma = _lib.as_f8(cos_gammaArray, 1)
    energies = _lib.as_f8(energies, 1)
    n = int(sqrt_numPix)
    if theta.shape != (n, n) or phi.shape != (n, n):
        raise ValueError("theta/phi must be [sqrt_numPix, sqrt_numPix]")
    atm = _lib.Atmosphere.get(atmosphere)
    flux = np.zeros(energies.shape[0], dtype=np.float64)
    rc = _lib.lib.xpsi_b200_integrate_time_invariance(
        float(R), float(omega), float(r_s), float(inclination), n, float(cellArea), _lib.dptr(radial),
        _lib.dptr(rsr), _lib.dptr(theta), _lib.dptr(phi), _lib.dptr(par), par.shape[2], int(numRays),
        _lib.dptr(deflection), _lib.dptr(cos_alpha), _lib.dptr(maxDeflection), _lib.dptr(cos_gamma),
        energies.shape[0], _lib.dptr(energies), atm.handle if atm is not None else None, int(atm_ext),
        int(image_order_limit) if image_order_limit is not None else 0, _lib.dptr(flux))
    if rc == _lib.ENUMERICAL:
        return (1, None)
    if rc == _lib.EUNSUPPORTED:
        raise NotImplementedError("xpsi_b200: " + _lib.last_error())
    _lib.check(rc)
    return (0, flux)
