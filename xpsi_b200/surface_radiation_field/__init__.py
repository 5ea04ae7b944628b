"""GPU drop-in for ``xpsi.surface_radiation_field`` (the point-wise intensity seam)."""
import numpy as np

from .. import _lib

_ATM = {"BB": 1, "Num4D": 2}      # XPSI_B200_ATM_BB / XPSI_B200_ATM_NUM4D


def intensity(energies, mu, local_variables, atmosphere=None, stokesQ=0, region_extension='hot',
              atmos_extension="BB", beam_opt=0, numTHREADS=1):
    """Same signature and units as xpsi/surface_radiation_field/core.pyx:125-308: photon specific
    intensity [photons/s/keV/cm^2/sr] at each (energy, mu, local-variables row).

    Covered: ``atmos_extension`` 'BB' and 'Num4D', ``region_extension`` 'hot' (``beam_opt`` 0-3) and
    'elsewhere'.  Polarimetry (``stokesQ=1``) and the burst / Num2D / user extensions raise
    ``NotImplementedError``.  ``numTHREADS`` is accepted and ignored.
    """
    if region_extension not in ('hot', 'elsewhere'):
        raise ValueError("Region extension module must be 'hot' or 'elsewhere'.")
    if stokesQ == 1:
        if region_extension == 'elsewhere':
            raise ValueError("StokesQ option is not allowed for the elsewhere extension.")
        raise NotImplementedError("xpsi_b200: Stokes Q intensities are not covered")
    if atmos_extension not in ("BB", "Num4D", "Pol_BB_Burst", "Pol_Num2D", "user"):
        raise ValueError("Atmosphere extension module must be 'BB', 'Num4D', 'Pol_BB_Burst', 'Pol_Num2D', or 'user'.")
    if atmos_extension not in _ATM:
        raise NotImplementedError("xpsi_b200: atmosphere extension %r is not covered" % atmos_extension)
    if atmos_extension == "Num4D" and atmosphere is None:
        raise ValueError("Atmosphere data must be loaded if using numerical atmosphere extension.")
    energies = _lib.as_f8(energies, 1)
    mu = _lib.as_f8(mu, 1)
    local_variables = _lib.as_f8(local_variables, 2)
    n = energies.shape[0]
    if mu.shape[0] != n or local_variables.shape[0] != n:
        raise ValueError("energies, mu and local_variables must have the same number of points")
    atm = _lib.Atmosphere.get(atmosphere) if atmos_extension == "Num4D" else None
    out = np.zeros(n, dtype=np.float64)
    rc = _lib.lib.xpsi_b200_intensity(n, _lib.dptr(energies), _lib.dptr(mu), _lib.dptr(local_variables),
                                      local_variables.shape[1], atm.handle if atm is not None else None,
                                      0 if region_extension == 'hot' else 1, _ATM[atmos_extension], int(beam_opt),
                                      _lib.dptr(out))
    _lib.check(rc)
    return out
