"""GPU drop-in for ``xpsi.cellmesh.integrator`` (the integrator without azimuthal invariance)."""
from .. import _lib
from .integrator_for_azimuthal_invariance import _integrate


def integrate(numThreads, R, omega, r_s, inclination, cellArea, radialCoords_of_parallels,
              r_s_over_r, theta, phi, srcCellParams, CELL_RADIATES, correction_srcCellParams,
              numRays, deflection, cos_alpha, lag, maxDeflection, cos_gammaArray, energies, leaves,
              phases, hot_atmosphere, elsewhere_atmosphere, hot_atm_ext, else_atm_ext, beam_opt,
              image_order_limit=None, R_in=1e6):
    """Same positional signature and return convention as xpsi/cellmesh/integrator.pyx:48-76
    (bound by ``HotRegion.symmetry = False``, xpsi/HotRegion.py:567-569, and by
    ``Everywhere(time_invariant=False)``, xpsi/Everywhere.py:330-332): ``(0, flux[N_E, N_P])`` on success,
    ``(1, None)`` on a numerical error.  The atmosphere is evaluated with every cell's own
    ``srcCellParams`` row.  ``numThreads`` is accepted and ignored; ``R_in`` is ignored as in the reference.
    """
    return _integrate(_lib.lib.xpsi_b200_integrate_general, R, omega, r_s, inclination, cellArea,
                      radialCoords_of_parallels, r_s_over_r, theta, phi, srcCellParams, CELL_RADIATES,
                      correction_srcCellParams, numRays, deflection, cos_alpha, lag, maxDeflection,
                      cos_gammaArray, energies, leaves, phases, hot_atmosphere, elsewhere_atmosphere,
                      hot_atm_ext, else_atm_ext, beam_opt, image_order_limit, R_in)
