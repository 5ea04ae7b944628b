"""Batched likelihood pipeline: integrate -> energy-integrate -> fold -> marginal
likelihood for B parameter vectors per call, entirely on one GPU.

This is the batched sibling of ``xpsi.Likelihood.__call__``
(xpsi/Likelihood.py:297-511): the scalar call evaluates one parameter vector
through Python glue between four compiled stages; here the same four stages run
as kernels over a leading batch axis and only ``lnL[B]`` / ``status[B]`` return.
Inputs are the per-member integrator arguments that ``HotRegion.embed``
produces (xpsi/HotRegion.py:1033-1070).
"""
import ctypes as C

import numpy as np

from . import _lib

MEMBER_FIELDS = ("cellArea", "phi", "theta", "radial", "r_s_over_r", "srcParams",
                 "deflection", "cos_alpha", "lag", "maxDeflection", "cos_gamma")


class HostBatch:
    """Padded, C-contiguous host arrays for B parameter vectors x M members."""

    def __init__(self, B, M, C_, max_rings, max_azi, n_rays, n_params, pinned=False):
        self.B, self.M = B, M
        Q = B * M
        self._pins = []
        alloc = self._pinned if pinned else np.zeros
        self.omega = alloc((B,), np.float64)
        self.inclination = alloc((B,), np.float64)
        self.d_sq = alloc((B,), np.float64)
        self.phase_shifts = alloc((B, C_), np.float64)
        self.n_rings = alloc((Q,), np.int32)
        self.n_azi = alloc((Q,), np.int32)
        self.cellArea = alloc((Q, max_rings, max_azi), np.float64)
        self.phi = alloc((Q, max_rings, max_azi), np.float64)
        self.theta = alloc((Q, max_rings), np.float64)
        self.radial = alloc((Q, max_rings), np.float64)
        self.r_s_over_r = alloc((Q, max_rings), np.float64)
        self.srcParams = alloc((Q, max_rings, n_params), np.float64)
        self.deflection = alloc((Q, max_rings, n_rays), np.float64)
        self.cos_alpha = alloc((Q, max_rings, n_rays), np.float64)
        self.lag = alloc((Q, max_rings, n_rays), np.float64)
        self.maxDeflection = alloc((Q, max_rings), np.float64)
        self.cos_gamma = alloc((Q, max_rings), np.float64)

    def _pinned(self, shape, dtype):
        import torch   # device-memory plumbing only: page-locked host staging
        t = torch.zeros(shape, dtype=torch.float64 if dtype == np.float64 else torch.int32).pin_memory()
        self._pins.append(t)           # the numpy view below borrows the tensor's storage; freed with the batch
        return t.numpy()

    def set_member(self, b, m, cellArea, theta, phi, radial, r_s_over_r, srcCellParams,
                   deflection, cos_alpha, lag, maxDeflection, cos_gamma):
        """Fill member ``m`` of parameter vector ``b`` from the reference-shaped
        integrator arguments (cellArea/theta/phi ``[R,A]``, srcCellParams ``[R,A,n]``)."""
        q = b * self.M + m
        R, A = cellArea.shape
        if R > self.cellArea.shape[1] or A > self.cellArea.shape[2]:
            raise ValueError("mesh %dx%d exceeds the padded size" % (R, A))
        self.n_rings[q] = R
        self.n_azi[q] = A
        self.cellArea[q].fill(0.0)
        self.cellArea[q, :R, :A] = cellArea
        self.phi[q, :R, :A] = phi
        self.theta[q, :R] = theta[:, 0]
        self.radial[q, :R] = radial
        self.r_s_over_r[q, :R] = r_s_over_r
        # all cells of a ring share the parameter vector of its first radiating
        # cell (integrator_for_azimuthal_invariance.pyx:286-296,463)
        rad = cellArea > 0.0
        J = np.argmax(rad, axis=1)
        self.srcParams[q, :R] = srcCellParams[np.arange(R), J]
        self.deflection[q, :R] = deflection
        self.cos_alpha[q, :R] = cos_alpha
        self.lag[q, :R] = lag
        self.maxDeflection[q, :R] = maxDeflection
        self.cos_gamma[q, :R] = cos_gamma

    def nbytes(self):
        return sum(getattr(self, f).nbytes for f in
                   ("omega", "inclination", "d_sq", "phase_shifts", "n_rings", "n_azi") + MEMBER_FIELDS)

    def struct(self):
        s = _lib.Batch()
        for f in ("omega", "inclination", "d_sq", "phase_shifts") + MEMBER_FIELDS:
            setattr(s, f, _lib.dptr(getattr(self, f)))
        s.n_rings = _lib.iptr(self.n_rings)
        s.n_azi = _lib.iptr(self.n_azi)
        return s


class SpotBatch:
    """Parameter-level inputs for B parameter vectors whose M hot-region members are simple circular spots.

    Per theta: the derived spacetime scalars of ``xpsi.Spacetime`` (xpsi/Spacetime.py:110-188); per member:
    colatitude, angular radius, log10 temperature and the azimuth offset (pi for an antiphased region,
    xpsi/HotRegion.py:834-837).  The mesh and rays are then built on the GPU.
    """

    def __init__(self, B, M, C_, mode_frequency, num_cells=1024, min_sqrt_num_cells=10, max_sqrt_num_cells=64):
        self.B, self.M = B, M
        z = lambda *shape: np.zeros(shape, dtype=np.float64)
        self.R_eq, self.r_s, self.epsilon, self.zeta = z(B), z(B), z(B), z(B)
        self.omega, self.inclination, self.d_sq = z(B), z(B), z(B)
        self.phase_shifts = z(B, C_)
        self.colatitude, self.ang_radius, self.temperature, self.phi_shift = z(B, M), z(B, M), z(B, M), z(B, M)
        self.mode_frequency = float(mode_frequency)
        self.num_cells, self.min_sqrt, self.max_sqrt = int(num_cells), int(min_sqrt_num_cells), int(max_sqrt_num_cells)
        # optional: masking region of each member and superseding / ceding pairing (see set_region)
        self.hole_radius = self.hole_colatitude = self.hole_azimuth = None
        self.partner = self.is_cede = None
        # optional [B, M, n_params-2]: local variables after (log T, log g), e.g. beaming parameters
        self.extra_params = None
        # optional int [M, 3]: (num_cells, min_sqrt_num_cells, max_sqrt_num_cells) of the hot region each member
        # belongs to, when the regions were constructed with different resolutions
        self.member_cells = None

    def set_region(self, super_member, cede_member=None, *, super_colatitude, super_radius, super_temperature,
                   omit_colatitude=None, omit_radius=None, omit_azimuth=None, cede_colatitude=None,
                   cede_radius=None, cede_azimuth=None, cede_temperature=None, is_antiphased=False):
        """Fill the members of one hot region from ``xpsi.HotRegion`` parameter values (arrays over the batch),
        the way ``HotRegion.embed`` hands them to the mesh routines (xpsi/HotRegion.py:819-865): the superseding
        member is masked by the omission region, the ceding member by the superseding region."""
        if self.hole_radius is None:
            z = lambda: np.zeros((self.B, self.M), dtype=np.float64)
            self.hole_radius, self.hole_colatitude, self.hole_azimuth = z(), z(), z()
            self.hole_colatitude[:] = self.colatitude
        pi_shift = np.pi if is_antiphased else 0.0
        m = super_member
        self.colatitude[:, m], self.ang_radius[:, m], self.temperature[:, m] = super_colatitude, super_radius, super_temperature
        self.hole_radius[:, m] = 0.0 if omit_radius is None else omit_radius
        self.hole_colatitude[:, m] = super_colatitude if omit_colatitude is None else omit_colatitude
        self.hole_azimuth[:, m] = 0.0 if omit_azimuth is None else omit_azimuth
        self.phi_shift[:, m] = -self.hole_azimuth[:, m] + pi_shift            # HotRegion.py:834-837
        if cede_member is not None:
            if self.partner is None:
                self.partner = -np.ones(self.M, dtype=np.int32)
                self.is_cede = np.zeros(self.M, dtype=np.int32)
            c = cede_member
            self.partner[m], self.partner[c], self.is_cede[c] = c, m, 1
            self.colatitude[:, c], self.ang_radius[:, c], self.temperature[:, c] = cede_colatitude, cede_radius, cede_temperature
            azi = np.zeros(self.B) if cede_azimuth is None else np.asarray(cede_azimuth, dtype=np.float64)
            self.hole_radius[:, c], self.hole_colatitude[:, c] = super_radius, super_colatitude
            self.hole_azimuth[:, c] = -azi                                     # HotRegion.py:865
            self.phi_shift[:, c] = azi + pi_shift                              # HotRegion.py:867-870
            self.hole_radius[:, m] = 0.0        # "omit=True and cede=True ignores the omission" (HotRegion.py:44)

    def set_spacetime(self, mass, radius, distance, cos_inclination, frequency):
        """Vectorised ``xpsi.Spacetime`` derived quantities (xpsi/Spacetime.py:110-188)."""
        from . import synthetic as syn
        mass, radius, distance, cos_i = [np.asarray(v, dtype=np.float64) for v in (mass, radius, distance, cos_inclination)]
        r_g = mass * syn.GM_SUN
        self.r_s[:] = 2.0 * r_g
        self.R_eq[:] = radius * syn.KM
        M = mass * syn.GM_SUN * syn.C_LIGHT * syn.C_LIGHT / syn.G_NEWTON
        Omega = 2.0 * np.pi * frequency
        self.omega[:] = Omega
        self.inclination[:] = np.arccos(cos_i)
        self.d_sq[:] = (distance * syn.KPC) ** 2
        self.zeta[:] = r_g / self.R_eq
        self.epsilon[:] = Omega ** 2 * self.R_eq ** 3 / (syn.G_NEWTON * M)

    def struct(self):
        s = _lib.SpotBatch()
        for f in ("R_eq", "r_s", "epsilon", "zeta", "omega", "inclination", "d_sq", "phase_shifts",
                  "colatitude", "ang_radius", "temperature", "phi_shift"):
            a = getattr(self, f)
            if not a.flags.c_contiguous:
                raise ValueError(f + " must be C-contiguous")
            setattr(s, f, _lib.dptr(a))
        s.mode_frequency = self.mode_frequency
        s.num_cells, s.min_sqrt_num_cells, s.max_sqrt_num_cells = self.num_cells, self.min_sqrt, self.max_sqrt
        if self.hole_radius is not None:
            for f in ("hole_radius", "hole_colatitude", "hole_azimuth"):
                setattr(s, f, _lib.dptr(getattr(self, f)))
        if self.partner is not None:
            s.partner, s.is_cede = _lib.iptr(self.partner), _lib.iptr(self.is_cede)
        if self.extra_params is not None:
            self.extra_params = np.ascontiguousarray(self.extra_params, dtype=np.float64)
            if self.extra_params.shape[:2] != (self.B, self.M):
                raise ValueError("extra_params must have shape [B, M, n_params-2]")
            s.extra_params = _lib.dptr(self.extra_params)
        if self.member_cells is not None:
            self.member_cells = np.ascontiguousarray(self.member_cells, dtype=np.int32)
            if self.member_cells.shape != (self.M, 3):
                raise ValueError("member_cells must have shape [M, 3]")
            s.member_cells = _lib.iptr(self.member_cells)
        return s


class _DeviceArray:
    """A device buffer owned by the pipeline, exposed through ``__cuda_array_interface__`` (version 3)."""

    def __init__(self, ptr, n, typestr, owner):
        self._owner = owner
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr=typestr, data=(int(ptr), False), version=3,
                                             strides=None, stream=None)   # ordered by the caller
        # (the producer is the library's stream, ``xpsi_b200_stream()``: consume under that stream)


class BatchedLikelihood:
    """Device-resident likelihood for a fixed model configuration.

    Parameters mirror what the reference objects hold: ``energies`` (keV,
    ``Signal.energies``), ``leaves``/``phases`` (radians, ``HotRegion``),
    ``response[n_chan, n_in]`` + ``energy_edges`` (``Instrument``), ``counts``,
    ``data_phases``, ``exposure_time`` (``Data``), ``support`` and the
    ``eval_marginal_likelihood`` settings (``CustomSignal``).
    """

    def __init__(self, *, member_component=(0,), max_rings=0, max_azi=0, n_rays=0, energies, leaves=None, phases=None,
                 hot_atm_ext=1, hot_atmosphere=None, image_order_limit=None, response, energy_edges,
                 counts, data_phases, exposure_time, support=None, epsilon=1.0e-3, sigmas=10.0,
                 llzero=-1.0e90, slim=20.0, allow_negative=False, n_params=2, max_batch=64,
                 phase_interpolant='Akima'):
        self._keep = []

        def keep(a, dt=np.float64):
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return a
        mc = keep(member_component, np.int32)
        self.n_members = int(mc.size)
        self.n_components = int(mc.max()) + 1
        # ``phases=None`` (or one phase): the star is an ``Everywhere(time_invariant=True)`` surface
        # (xpsi/Everywhere.py:577-601) -- no hot regions, one phase column; its mesh settings come with
        # ``set_extras(everywhere=...)``
        self.time_invariant = phases is None or np.size(phases) == 1
        if self.time_invariant:
            if self.n_members != 1:
                raise ValueError("a time-invariant (Everywhere) pipeline has exactly one member")
            phases, leaves = np.zeros(1), np.zeros(1)
        energies, leaves, phases = keep(energies), keep(leaves), keep(phases)
        response, energy_edges = keep(response), keep(energy_edges)
        counts, data_phases = keep(counts), keep(data_phases)
        n_chan, n_in = response.shape
        if support is None:
            support = -1.0 * np.ones((n_chan, 2))
            support[:, 0] = 0.0
        support = keep(support)
        self.atm = _lib.Atmosphere.get(hot_atmosphere) if hot_atm_ext == 2 else None
        cfg = _lib.PipelineConfig()
        cfg.n_components, cfg.n_members, cfg.member_component = self.n_components, self.n_members, _lib.iptr(mc)
        cfg.max_rings, cfg.max_azi, cfg.n_rays, cfg.n_params = int(max_rings), int(max_azi), int(n_rays), int(n_params)
        cfg.n_energies, cfg.energies = energies.size, _lib.dptr(energies)
        cfg.n_leaves, cfg.leaves = (0 if self.time_invariant else leaves.size), _lib.dptr(leaves)
        cfg.n_phases, cfg.phases = phases.size, _lib.dptr(phases)
        cfg.hot_atm_ext = int(hot_atm_ext)
        cfg.hot_atmosphere = self.atm.handle if self.atm is not None else None
        cfg.image_order_limit = int(image_order_limit) if image_order_limit else 0
        cfg.phase_interpolant = _lib.INTERPOLANTS[phase_interpolant]
        cfg.n_in, cfg.energy_edges = n_in, _lib.dptr(energy_edges)
        cfg.n_chan, cfg.response = n_chan, _lib.dptr(response)
        cfg.n_bins, cfg.data_phases = data_phases.size - 1, _lib.dptr(data_phases)
        cfg.counts, cfg.support = _lib.dptr(counts), _lib.dptr(support)
        cfg.exposure_time, cfg.epsilon, cfg.sigmas = float(exposure_time), float(epsilon), float(sigmas)
        cfg.llzero, cfg.slim, cfg.allow_negative = float(llzero), float(slim), int(bool(allow_negative))
        self.cfg = cfg
        self.max_batch = int(max_batch)
        self.shape = dict(max_rings=int(max_rings), max_azi=int(max_azi), n_rays=int(n_rays),
                          n_params=int(n_params), n_energies=energies.size, n_phases=phases.size,
                          n_chan=n_chan, n_in=n_in, n_bins=data_phases.size - 1)
        self.handle = _lib.lib.xpsi_b200_pipeline_create(C.byref(cfg), self.max_batch)
        if not self.handle:
            raise _lib.XpsiB200Error("pipeline_create failed: %s" % _lib.last_error())
        self.signals = [dict(n_chan=n_chan, n_in=n_in, n_bins=data_phases.size - 1)]

    def add_signal(self, *, response, energy_edges, counts, data_phases, exposure_time, support=None, epsilon=1.0e-3,
                   sigmas=10.0, llzero=-1.0e90, slim=20.0, allow_negative=False, attenuation=None):
        """Register one more (instrument, data) pair behind the same integrator stage -- a second ``xpsi.Signal``
        of the photosphere (xpsi/Likelihood.py:346-420; docs/source/Instrument_synergy.ipynb).  The photosphere
        signal is computed once at the pipeline's energies (the reference gives all signals of a photosphere one
        energy array, Likelihood.py:102-107); this signal gets its own energy integration, interstellar
        attenuation (``attenuation[n_in]`` at unit power, raised to the batch's ``att_power``), response fold and
        background-marginalised likelihood, and the joint log-likelihood is the sum (Likelihood.py:494-500).
        Returns the signal's index (the constructor's signal is 0)."""
        response = np.ascontiguousarray(response, dtype=np.float64)
        n_chan, n_in = response.shape
        energy_edges = np.ascontiguousarray(energy_edges, dtype=np.float64)
        counts = np.ascontiguousarray(counts, dtype=np.float64)
        data_phases = np.ascontiguousarray(data_phases, dtype=np.float64)
        if energy_edges.shape != (n_in + 1,) or counts.shape != (n_chan, data_phases.size - 1):
            raise ValueError("add_signal: response [n_chan, n_in], energy_edges [n_in+1], counts [n_chan, n_bins] "
                             "and data_phases [n_bins+1] do not fit together")
        if support is None:
            support = -1.0 * np.ones((n_chan, 2))
            support[:, 0] = 0.0
        support = np.ascontiguousarray(support, dtype=np.float64)
        sc = _lib.SignalConfig()
        sc.n_in, sc.energy_edges = n_in, _lib.dptr(energy_edges)
        sc.n_chan, sc.response = n_chan, _lib.dptr(response)
        sc.n_bins, sc.data_phases = data_phases.size - 1, _lib.dptr(data_phases)
        sc.counts, sc.support = _lib.dptr(counts), _lib.dptr(support)
        sc.exposure_time, sc.epsilon, sc.sigmas = float(exposure_time), float(epsilon), float(sigmas)
        sc.llzero, sc.slim, sc.allow_negative = float(llzero), float(slim), int(bool(allow_negative))
        if attenuation is not None:
            att = np.ascontiguousarray(attenuation, dtype=np.float64)
            if att.shape != (n_in,):
                raise ValueError("one attenuation factor per instrument input interval is required")
            sc.attenuation = _lib.dptr(att)
        idx = _lib.lib.xpsi_b200_pipeline_add_signal(self.handle, C.byref(sc))
        if idx < 0:
            _lib.check(idx)
        self.signals.append(dict(n_chan=n_chan, n_in=n_in, n_bins=data_phases.size - 1))
        return idx

    def upload_signal_shifts(self, B, shifts):
        """``shifts[B, n_signals]`` in cycles: each signal's own phase shift, added to the hot regions' shifts for
        that signal (``Signal.shifts``, xpsi/Signal.py:581-583); ``None`` switches them off."""
        if shifts is None:
            _lib.check(_lib.lib.xpsi_b200_pipeline_upload_signal_shifts(self.handle, B, None))
            return
        a = np.ascontiguousarray(shifts, dtype=np.float64)
        if a.shape != (B, len(self.signals)):
            raise ValueError("signal shifts must have shape (%d, %d)" % (B, len(self.signals)))
        _lib.check(_lib.lib.xpsi_b200_pipeline_upload_signal_shifts(self.handle, B, _lib.dptr(a)))

    def fetch_signal(self, signal, B):
        """``(folded[B, C, n_chan, n_phases], expected[B, n_chan, n_bins], lnL[B])`` of one signal."""
        g = self.signals[signal]
        f = np.empty((B, self.n_components, g["n_chan"], self.shape["n_phases"]))
        e = np.empty((B, g["n_chan"], g["n_bins"]))
        l = np.empty(B)
        _lib.check(_lib.lib.xpsi_b200_pipeline_fetch_signal(self.handle, signal, B, _lib.dptr(f), _lib.dptr(e),
                                                            _lib.dptr(l)))
        return f, e, l

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib.xpsi_b200_pipeline_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def set_deterministic(self, on=True):
        """Ring sums by an ordered two-stage reduction instead of fp64 atomics: lnL becomes bitwise reproducible
        (run to run, and whatever the position of a parameter vector in its batch)."""
        _lib.check(_lib.lib.xpsi_b200_pipeline_set_deterministic(self.handle, int(bool(on))))

    def set_extras(self, elsewhere=None, attenuation=None, beam_opt=0, everywhere=None):
        """Optional model components (set once).

        ``everywhere`` (time-invariant pipelines only): the same dict as ``elsewhere`` -- the closed equal-area
        surface mesh of ``xpsi.Everywhere`` (xpsi/Everywhere.py:117-123 for the defaults) with a uniform
        temperature; the spectrum IS the star's signal (xpsi/Photosphere.py:560-566), one phase column.

        ``elsewhere``: dict(sqrt_num_cells, num_rays, atm_ext, atmosphere=None, image_order_limit=None) adds
        ``xpsi.Elsewhere`` (xpsi/Elsewhere.py:129-134 for the defaults the reference uses): its spectrum is
        added to member 0 (xpsi/Photosphere.py:589-592) and the hot members get the elsewhere correction.
        ``attenuation``: interstellar attenuation factor per instrument input interval at unit power
        (xpsi/Interstellar.py:27-58); the per-theta power comes with ``upload_extras(att_power=...)``.
        ``beam_opt``: beaming option of the hot regions (xpsi/HotRegion.py:184-200)."""
        x = _lib.PipelineExtras()
        self._extras_keep = []
        if everywhere is not None:
            if not self.time_invariant or elsewhere is not None:
                raise ValueError("everywhere= belongs to a time-invariant pipeline (phases=None) without elsewhere")
            elsewhere = everywhere
        elif self.time_invariant and elsewhere is not None:
            raise ValueError("a time-invariant pipeline takes everywhere=, not elsewhere=")
        if elsewhere is not None:
            x.elsewhere = 1
            x.else_sqrt_num_cells = int(elsewhere["sqrt_num_cells"])
            x.else_num_rays = int(elsewhere["num_rays"])
            x.else_atm_ext = int(elsewhere["atm_ext"])
            x.else_image_order_limit = int(elsewhere.get("image_order_limit") or 0)
            if x.else_atm_ext == 2:
                atm = _lib.Atmosphere.get(elsewhere["atmosphere"])
                self._extras_keep.append(atm)
                x.elsewhere_atmosphere = atm.handle
            self.elsewhere = dict(n=x.else_sqrt_num_cells, n_rays=x.else_num_rays)
        if attenuation is not None:
            att = np.ascontiguousarray(attenuation, dtype=np.float64)
            if att.shape != (self.shape["n_in"],):
                raise ValueError("one attenuation factor per instrument input interval is required")
            self._extras_keep.append(att)
            x.attenuation = _lib.dptr(att)
        x.beam_opt = int(beam_opt)
        _lib.check(_lib.lib.xpsi_b200_pipeline_set_extras(self.handle, C.byref(x)))
        self.shape["has_elsewhere"] = elsewhere is not None
        self.shape["has_attenuation"] = attenuation is not None

    def upload_extras(self, B, att_power=None, else_temperature=None, elsewhere=None, correction_srcParams=None):
        """Per-batch inputs of the optional components for the next evaluation.

        theta-level path: ``else_temperature[B]`` (+ ``att_power[B]``).  Mesh-level path: ``elsewhere`` = dict of
        the arrays ``Elsewhere.embed`` produced with a leading batch axis (cellArea[B], radial, r_s_over_r,
        maxDeflection, cos_gamma [B,n], theta, phi [B,n,n], srcParams [B,n,n,2], deflection, cos_alpha
        [B,n,num_rays]) and ``correction_srcParams[B*M, max_rings, n_params]``."""
        x = _lib.BatchExtras()
        keep = []

        def put(field, arr, shape=None):
            a = np.ascontiguousarray(arr, dtype=np.float64)
            if shape is not None and a.shape != shape:
                raise ValueError("%s must have shape %r, got %r" % (field, shape, a.shape))
            keep.append(a)
            setattr(x, field, _lib.dptr(a))
        if att_power is not None:
            put("att_power", att_power, (B,))
        if else_temperature is not None:
            put("else_temperature", else_temperature, (B,))
        if elsewhere is not None:
            n, nr = self.elsewhere["n"], self.elsewhere["n_rays"]
            shapes = dict(cellArea=(B,), radial=(B, n), r_s_over_r=(B, n), theta=(B, n, n), phi=(B, n, n),
                          srcParams=(B, n, n, 2), deflection=(B, n, nr), cos_alpha=(B, n, nr),
                          maxDeflection=(B, n), cos_gamma=(B, n))
            for k, shp in shapes.items():
                put("else_" + k, elsewhere[k], shp)
        if correction_srcParams is not None:
            s = self.shape
            put("correction_srcParams", correction_srcParams, (B * self.n_members, s["max_rings"], s["n_params"]))
        _lib.check(_lib.lib.xpsi_b200_pipeline_upload_extras(self.handle, B, C.byref(x)))

    def fetch_elsewhere(self, B):
        out = np.empty((B, self.shape["n_energies"]))
        _lib.check(_lib.lib.xpsi_b200_pipeline_fetch_elsewhere(self.handle, B, _lib.dptr(out)))
        return out

    def new_batch(self, B, pinned=False):
        s = self.shape
        return HostBatch(B, self.n_members, self.n_components, s["max_rings"], s["max_azi"],
                         s["n_rays"], s["n_params"], pinned=pinned)

    def __call__(self, batch):
        """Host buffers in, ``(lnL[B], status[B])`` out (H2D + kernels + D2H)."""
        lnL = np.empty(batch.B, dtype=np.float64)
        status = np.empty(batch.B, dtype=np.int32)
        st = batch.struct()
        _lib.check(_lib.lib.xpsi_b200_pipeline_eval(self.handle, batch.B, C.byref(st),
                                                    _lib.dptr(lnL), _lib.iptr(status)))
        return lnL, status

    def new_spot_batch(self, B, mode_frequency, **kw):
        sb = SpotBatch(B, self.n_members, self.n_components, mode_frequency, **kw)
        if not self.time_invariant and sb.max_sqrt > min(self.shape["max_rings"], self.shape["max_azi"]):
            raise ValueError("max_sqrt_num_cells = %d exceeds the pipeline's padded mesh (%d x %d): the embed "
                             "would refuse every parameter vector that allocates more rings"
                             % (sb.max_sqrt, self.shape["max_rings"], self.shape["max_azi"]))
        return sb

    def eval_spots(self, spots):
        """theta-level call: embed (mesh + rays) on the GPU, then the four likelihood stages."""
        lnL = np.empty(spots.B, dtype=np.float64)
        status = np.empty(spots.B, dtype=np.int32)
        st = spots.struct()
        _lib.check(_lib.lib.xpsi_b200_pipeline_eval_spots(self.handle, spots.B, C.byref(st), _lib.dptr(lnL),
                                                          _lib.iptr(status)))
        return lnL, status

    def embed_spots(self, spots):
        st = spots.struct()
        _lib.check(_lib.lib.xpsi_b200_pipeline_embed_spots(self.handle, spots.B, C.byref(st)))

    def fetch_embed(self, B):
        """Integrator inputs produced by the last embed, as a dict of padded arrays."""
        s = self.shape
        Q, R, A, NR = B * self.n_members, s["max_rings"], s["max_azi"], s["n_rays"]
        out = dict(n_rings=np.empty(Q, np.int32), cellArea=np.empty((Q, R, A)), phi=np.empty((Q, R, A)),
                   theta=np.empty((Q, R)), radial=np.empty((Q, R)), srcParams=np.empty((Q, R, s["n_params"])),
                   cos_gamma=np.empty((Q, R)), deflection=np.empty((Q, R, NR)), cos_alpha=np.empty((Q, R, NR)),
                   lag=np.empty((Q, R, NR)), maxDeflection=np.empty((Q, R)))
        _lib.check(_lib.lib.xpsi_b200_pipeline_fetch_embed(
            self.handle, B, _lib.iptr(out["n_rings"]), *[_lib.dptr(out[k]) for k in
                                                         ("cellArea", "phi", "theta", "radial", "srcParams", "cos_gamma",
                                                          "deflection", "cos_alpha", "lag", "maxDeflection")]))
        return out

    def upload(self, batch):
        st = batch.struct()
        _lib.check(_lib.lib.xpsi_b200_pipeline_upload(self.handle, batch.B, C.byref(st)))

    def eval_resident(self, B):
        _lib.check(_lib.lib.xpsi_b200_pipeline_eval_resident(self.handle, B))

    def download(self, B):
        lnL = np.empty(B, dtype=np.float64)
        status = np.empty(B, dtype=np.int32)
        _lib.check(_lib.lib.xpsi_b200_pipeline_download(self.handle, B, _lib.dptr(lnL), _lib.iptr(status)))
        return lnL, status

    def fetch(self, B, flux=True, folded=True, expected=True):
        s = self.shape
        f = np.empty((B * self.n_members, s["n_energies"], s["n_phases"])) if flux else None
        g = np.empty((B, self.n_components, s["n_chan"], s["n_phases"])) if folded else None
        e = np.empty((B, s["n_chan"], s["n_bins"])) if expected else None
        _lib.check(_lib.lib.xpsi_b200_pipeline_fetch(
            self.handle, B, _lib.dptr(f) if flux else None, _lib.dptr(g) if folded else None,
            _lib.dptr(e) if expected else None))
        return f, g, e

    def count_work(self, enable=True):
        """Toggle the integrator's algorithmic-work counters; returns dict(H, V, RI, K) summed
        over the evaluations since counting was last enabled (enabling resets them)."""
        out = (C.c_ulonglong * 4)()
        _lib.check(_lib.lib.xpsi_b200_pipeline_work_counters(self.handle, int(enable), out))
        return dict(H=out[0], V=out[1], RI=out[2], K=out[3])

    # ---- sweep: N parameter vectors resident on the device, evaluated block by block -------------------
    def sweep_upload(self, spots, att_power=None, else_temperature=None, signal_shifts=None):
        """Upload the parameter-level inputs of ``spots.B`` parameter vectors once (any N; evaluation
        proceeds in blocks of ``max_batch``).  ``signal_shifts[N, n_signals]``: see ``upload_signal_shifts``."""
        st = spots.struct()
        keep = []

        def opt(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != (spots.B,):
                raise ValueError("per-theta extras must have shape (%d,)" % spots.B)
            keep.append(a)
            return _lib.dptr(a)
        _lib.check(_lib.lib.xpsi_b200_pipeline_sweep_upload(self.handle, spots.B, C.byref(st), opt(att_power),
                                                            opt(else_temperature)))
        self._sweep_n = spots.B
        if signal_shifts is not None:
            a = np.ascontiguousarray(signal_shifts, dtype=np.float64)
            if a.shape != (spots.B, len(self.signals)):
                raise ValueError("signal shifts must have shape (%d, %d)" % (spots.B, len(self.signals)))
            _lib.check(_lib.lib.xpsi_b200_pipeline_sweep_upload_signal_shifts(self.handle, spots.B, _lib.dptr(a)))

    def sweep_run(self, first=0, count=None):
        """Queue embed + the four stages for rows ``[first, first+count)`` of the uploaded sweep (no sync)."""
        count = self._sweep_n - first if count is None else count
        _lib.check(_lib.lib.xpsi_b200_pipeline_sweep_run(self.handle, first, count))

    def sweep_download(self, first=0, count=None):
        count = self._sweep_n - first if count is None else count
        lnL = np.empty(count, dtype=np.float64)
        status = np.empty(count, dtype=np.int32)
        _lib.check(_lib.lib.xpsi_b200_pipeline_sweep_download(self.handle, first, count, _lib.dptr(lnL),
                                                              _lib.iptr(status)))
        return lnL, status

    def sweep_device_results(self):
        """Device arrays ``lnL[N]`` (float64) and ``status[N]`` (int32) of the sweep as objects exposing
        ``__cuda_array_interface__`` -- what a collective reads in place (``torch.as_tensor(obj, device=...)``)."""
        a, b = C.c_void_p(), C.c_void_p()
        _lib.check(_lib.lib.xpsi_b200_pipeline_sweep_results(self.handle, C.byref(a), C.byref(b)))
        return (_DeviceArray(a.value, self._sweep_n, "<f8", self), _DeviceArray(b.value, self._sweep_n, "<i4", self))

    def sweep_spots(self, spots, att_power=None, else_temperature=None, signal_shifts=None):
        """Upload once, evaluate every block, download once: ``(lnL[N], status[N])``."""
        self.sweep_upload(spots, att_power, else_temperature, signal_shifts)
        self.sweep_run()
        return self.sweep_download()

    def eval_spots_resident(self, B):
        """embed + four stages on the spot batch already uploaded by ``embed_spots`` (kernels only)."""
        _lib.check(_lib.lib.xpsi_b200_pipeline_eval_spots_resident(self.handle, B))

    def stage_ms(self):
        ms = (C.c_float * 6)()
        _lib.check(_lib.lib.xpsi_b200_pipeline_stage_ms(self.handle, ms))
        return dict(embed=ms[4], integrate=ms[0], energy=ms[1], fold=ms[2], marginal=ms[3], flux_kernel=ms[5])
