"""Deterministic synthetic inputs for the likelihood hot path.

Nothing here reads the reference or the oracle: the NSX-shaped atmosphere
table, the NICER-like response and the parameter-vector prior are closed-form,
so tests, ``bench.py`` and the golden-vector generator all see the same
workload on any box.

Configurations
--------------
``C1``  examples_fast (ST, blackbody): exactly the settings of the reference's
        known-answer test (xpsi/tests/test_likelihood.py:15-134).
``M2``  ST-U with a 4-D NSX-shaped table (35,14,67,166), 2 hot regions, 32
        cells, 100 leaves/phases, 200 rays, 128 energies, NICER-like
        270-channel x 1500-input response, 32 phase bins
        (examples/examples_modeling_tutorial/TestRun_Num.py:121-175,199).
"""
import math

import numpy as np

# physical constants, xpsi/global_imports.py:70-80
C_LIGHT = 2.99792458e8
KPC = 3.08567758e19
KEV = 1.60217662e-16
GM_SUN = 0.5 * 2.95325024e3
G_NEWTON = 6.6740831e-11
H_KEV = 4.135667662e-18
KM = 1.0e3
K_B = 1.38064852e-23

NSX_SHAPE = (35, 14, 67, 166)


def nsx_like_table(shape=NSX_SHAPE):
    """Analytic stand-in for ``nsx_H_v200804.out`` with the real table shape.

    Returns ``(logT, logg, mu, logE, buf)`` in the layout the reference hands
    to ``init_preload`` (xpsi/surface_radiation_field/preload.pyx:6-44,
    xpsi/Photosphere.py:208-217): ascending axes, ``buf`` C-order [T,g,mu,E]
    holding I_nu / T^3 in cgs.  The profile is a limb-darkened, mildly
    gravity/temperature dependent Planck function -- smooth and positive, so
    cubic-Lagrange interpolation error is far below the parity tolerance.
    """
    nT, ng, nmu, nE = shape
    logT = np.linspace(5.1, 6.8, nT)
    logg = np.linspace(13.7, 15.0, ng)
    mu = np.linspace(1.0e-3, 1.0, nmu)
    logE = np.linspace(-1.3, 2.0, nE)
    x = 10.0 ** logE
    planck = 1.333e-16 * x ** 3 / np.expm1(x)                      # 2k^3/(h^2c^2) x^3/(e^x-1)
    beam = (0.4 + 0.6 * mu[:, None]) ** (1.0 + 0.2 * np.tanh(logE[None, :]))
    gfac = 1.0 + 0.03 * (logg - 14.3)
    tfac = 10.0 ** (0.1 * (logT - 6.0))
    buf = (tfac[:, None, None, None] * gfac[None, :, None, None]
           * (beam * planck[None, :])[None, None, :, :])
    return (np.ascontiguousarray(logT), np.ascontiguousarray(logg),
            np.ascontiguousarray(mu), np.ascontiguousarray(logE),
            np.ascontiguousarray(buf.reshape(-1)))


def nicer_like_response(n_in=1500, chan_lo=30, chan_hi=300):
    """Synthetic NICER-like RMFxARF (cm^2): Gaussian redistribution with a
    low-energy shelf, times a smooth ARF.

    Returns ``(matrix[n_chan, n_in], energy_edges[n_in+1], channels[n_chan],
    channel_edges[n_chan+1])`` -- the four arguments of ``xpsi.Instrument``
    (xpsi/Instrument.py:60-144).
    """
    edges = 0.2 + 0.005 * np.arange(n_in + 1)
    mids = 0.5 * (edges[:-1] + edges[1:])
    channels = np.arange(chan_lo, chan_hi)
    ch_edges = 0.01 * np.arange(chan_lo, chan_hi + 1)
    sigma = 0.02 + 0.04 * np.sqrt(mids)
    arf = 1800.0 * np.exp(-np.log(mids / 1.5) ** 2 / 0.8)
    erf = np.vectorize(math.erf)
    z_hi = (ch_edges[1:, None] - mids[None, :]) / (math.sqrt(2.0) * sigma[None, :])
    z_lo = (ch_edges[:-1, None] - mids[None, :]) / (math.sqrt(2.0) * sigma[None, :])
    rmf = 0.5 * (erf(z_hi) - erf(z_lo))
    rmf[rmf < 1.0e-12] = 0.0
    # low-energy redistribution shelf (partial charge collection), 0.2% total:
    # every input energy reaches some channel, as for a real RMF
    ch_mid = 0.5 * (ch_edges[1:] + ch_edges[:-1])
    shelf = (ch_mid[:, None] < mids[None, :]) * (2.0e-3 * 0.01 / mids[None, :])
    rmf = rmf + shelf
    matrix = np.ascontiguousarray(rmf * arf[None, :])
    return matrix, edges, channels, ch_edges


def c1_response():
    """Diagonal 1800 cm^2 response of xpsi/tests/test_likelihood.py:33-65.

    Returns ``(matrix[291,291], energy_edges[292], channels[291],
    channel_edges[292])`` for inputs 10..300 of the 0.1+0.005k keV grid.
    """
    lo = [0.1]
    while lo[-1] <= 14.995:
        lo.append(lo[-1] + 0.005)
    lo = np.array(lo)
    hi = lo + 0.005
    n = 291
    matrix = np.diag(np.full(n, 1800.0))
    edges = np.zeros(n + 1)
    edges[0] = lo[10]
    edges[1:] = hi[10:301]
    channels = np.arange(10, 301)
    ch_edges = np.arange(0, 15.01, 0.01)[10:302]
    return matrix, edges, channels, ch_edges


def signal_energies(matrix, energy_edges, num_energies):
    """``Signal`` energy grid: logspace over the first..last non-zero response
    columns (xpsi/Signal.py:324-354,1092-1197)."""
    nz = np.nonzero(matrix.sum(axis=0) > 0.0)[0]
    a, b = nz[0], nz[-1] + 1
    return np.logspace(np.log10(energy_edges[a]), np.log10(energy_edges[b]),
                       int(num_energies), base=10.0), (int(a), int(b))


class SpacetimeScalars:
    """Derived spacetime scalars, xpsi/Spacetime.py:110-188."""

    def __init__(self, mass, radius, distance, cos_inclination, frequency):
        self.mass = float(mass)
        self.radius = float(radius)
        self.distance = float(distance)
        self.cos_i = float(cos_inclination)
        self.frequency = float(frequency)
        self.r_g = self.mass * GM_SUN
        self.r_s = 2.0 * self.r_g
        self.R = self.radius * KM
        self.M = self.mass * GM_SUN * C_LIGHT * C_LIGHT / G_NEWTON
        self.Omega = 2.0 * math.pi * self.frequency
        self.i = math.acos(self.cos_i)
        self.d_sq = (self.distance * KPC) ** 2
        self.zeta = self.r_g / self.R
        self.epsilon = self.Omega ** 2 * self.R ** 3 / (G_NEWTON * self.M)
        self.R_r_s = self.R / self.r_s


# ST-U free-parameter order, TestRun_Num.py:301-311
M2_NAMES = ("mass", "radius", "distance", "cos_inclination",
            "p__phase_shift", "p__super_colatitude", "p__super_radius",
            "p__super_temperature",
            "s__phase_shift", "s__super_colatitude", "s__super_radius")

M2_TRUE = np.array([1.4, 12.0, 1.5, 0.5,
                    0.0, 1.0, 0.25, 6.55,
                    0.025, 2.2, 0.2])

M2_BOUNDS = np.array([[1.0, 2.0], [10.0, 14.0], [1.0, 2.5], [0.05, 0.95],
                      [-0.25, 0.75], [0.3, 1.4], [0.08, 0.45], [6.4, 6.7],
                      [-0.25, 0.75], [1.7, 2.8], [0.08, 0.45]])
M2_SECONDARY_DT = 0.2   # s__super_temperature = p__super_temperature - 0.2
M2_FREQUENCY = 300.0
M2_EXPOSURE = 1.0e6
M2_BACKGROUND_RATE = 1.0e-3   # counts/s/channel, flat


def m2_theta_batch(n, seed=20261017):
    """``n`` ST-U parameter vectors, uniform in ``M2_BOUNDS`` (in-table temperatures; the two spots cannot
    overlap inside the box; spots covering a pole are kept -- the embed meshes them as polar caps).  The list is
    a prefix-stable stream: row k is the same whatever ``n`` (11 uniforms per candidate, in order)."""
    rng = np.random.default_rng(seed)
    d = len(M2_NAMES)
    out = np.empty((n, d))
    k = 0
    while k < n:
        u = rng.uniform(size=(max(n - k, 16), d))
        th = M2_BOUNDS[:, 0] + u * (M2_BOUNDS[:, 1] - M2_BOUNDS[:, 0])
        # compactness: R >= 3 r_g  (TestRun_Num.py CustomPrior)
        th = th[th[:, 1] * KM >= 3.0 * th[:, 0] * GM_SUN]
        take = min(n - k, th.shape[0])
        out[k:k + take] = th[:take]
        k += take
    return out


def m2_prior():
    """The prior of the list above as a vectorised ``xpsi.Prior`` stand-in (xpsi_b200.sampling.BoxPrior)."""
    from .sampling import BoxPrior
    return BoxPrior(M2_NAMES, M2_BOUNDS, rules=(lambda P: P[:, 1] * KM >= 3.0 * P[:, 0] * GM_SUN,))


def m2_bench_thetas(first, count):
    """Rows ``[first, first+count)`` of the deterministic ST-U parameter-vector list that the bench, the
    sweep and the parity fixtures share: row 0 is ``M2_TRUE`` and row 1 is ``m2_theta_batch(4)[1]`` (the two
    parameter vectors of the golden fixture ``m2_stu_nsx.npz``), followed by ``m2_theta_batch`` draws."""
    head = np.vstack([M2_TRUE, m2_theta_batch(4)[1]])
    full = np.vstack([head, m2_theta_batch(first + count)])
    return np.ascontiguousarray(full[first:first + count])


def m2_near_truth_thetas(n, seed=7, scale=2.0e-3):
    """``n`` parameter vectors scattered around ``M2_TRUE`` (Gaussian, ``scale`` x the prior width per
    parameter): the region a converged sampler spends its time in, |lnL| ~ 4e4-1e5 for the fixture data."""
    rng = np.random.default_rng(seed)
    width = M2_BOUNDS[:, 1] - M2_BOUNDS[:, 0]
    return M2_TRUE[None, :] + scale * width[None, :] * rng.standard_normal((n, len(M2_NAMES)))


def m2_compact_thetas():
    """12 parameter vectors with stars close to the photon-sphere limit ``R = 3 r_g`` (mass 2.2, radius set
    from ``R / r_g``): six whose rays bend by more than 2 pi (three image orders are summed under
    ``image_order_limit = 3``) and six with two image orders.  The remaining parameters are rows 2..13 of the
    bench list.  Outside the prior box of ``M2_BOUNDS`` but inside the reference's strict parameter bounds."""
    x = np.array([3.005, 3.01, 3.02, 3.03, 3.035, 3.04, 3.06, 3.1, 3.15, 3.2, 3.3, 3.45])
    th = m2_bench_thetas(2, x.size)
    th[:, 0] = 2.2
    th[:, 1] = x * 2.2 * GM_SUN / KM
    return th


def m2_spot_batch(pipe, thetas):
    """Map ST-U parameter vectors (order ``M2_NAMES``) onto the pipeline's parameter-level inputs:
    two circular spots, the secondary antiphased with ``T_s = T_p - M2_SECONDARY_DT``
    (TestRun_Num.py:121-175)."""
    thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    B = thetas.shape[0]
    sb = pipe.new_spot_batch(B, M2_FREQUENCY, num_cells=1024, min_sqrt_num_cells=10, max_sqrt_num_cells=64)
    sb.set_spacetime(thetas[:, 0], thetas[:, 1], thetas[:, 2], thetas[:, 3], M2_FREQUENCY)
    sb.phase_shifts[:, 0] = thetas[:, 4]
    sb.phase_shifts[:, 1] = thetas[:, 8]
    sb.colatitude[:, 0], sb.ang_radius[:, 0], sb.temperature[:, 0] = thetas[:, 5], thetas[:, 6], thetas[:, 7]
    sb.colatitude[:, 1], sb.ang_radius[:, 1] = thetas[:, 9], thetas[:, 10]
    sb.temperature[:, 1] = thetas[:, 7] - M2_SECONDARY_DT
    sb.phi_shift[:, 0] = 0.0
    sb.phi_shift[:, 1] = math.pi
    return sb
