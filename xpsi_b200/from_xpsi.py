"""Build the GPU likelihood from an already constructed reference model: ``from_xpsi(xpsi.Likelihood(...))``.

A user of the reference has a tree of objects -- ``Star(Spacetime, Photosphere(HotRegion(s) | Elsewhere |
Everywhere))`` and one or more ``Signal(Data, Instrument, Interstellar)`` -- wrapped in ``xpsi.Likelihood(star,
signals, ...)`` (xpsi/Likelihood.py:61-160).  ``from_xpsi`` reads the *settings* of those objects once (resolution,
atmosphere tables, response, data, likelihood tolerances), creates the device-resident
:class:`~xpsi_b200.pipeline.BatchedLikelihood` (one integrator stage + one energy-integrate / fold / likelihood
stage per signal), and returns an :class:`xpsi_b200.likelihood.Likelihood` whose calls take parameter vectors in
the reference object's own order (``likelihood.names``):

    gpu = xpsi_b200.from_xpsi(likelihood)          # likelihood: the reference's object, unchanged
    gpu(p)                                         # == likelihood(p, force=True), xpsi/Likelihood.py:450-511
    gpu.batch(P)                                   # P[B, d] -> (lnL[B], status[B]) in blocks on the GPU

Parameter *values* are obtained through the reference's own parameter machinery: for every row the vector is
written with ``ParameterSubspace.__call__`` and each model object is asked for its parameters
(``hot['super_colatitude']`` ... -- xpsi/HotRegion.py:780-865 reads exactly these), so derived parameters
(``xpsi.Derive`` instances, e.g. a secondary temperature tied to the primary) and fixed values behave as in the
reference.  That walk is host Python (tens of microseconds per parameter vector); models that need the last
factor of throughput pass a vectorised ``fill`` to :class:`xpsi_b200.likelihood.Likelihood` instead.

Nothing here imports the reference: the objects are duck-typed, handed in by the caller.  Everything that is
computed is computed on the GPU; a model outside the kernels' coverage raises ``NotImplementedError`` at build time.

Covered: hot regions with azimuthal symmetry made of a superseding member with optional omission hole and
optional ceding member (ST / CST / EST / PST / CDT / EDT / PDT and their -U combinations), blackbody or Num4D
atmosphere, ``Elsewhere``, ``Everywhere(time_invariant=True)``, one or more signals with an optional
``Interstellar`` whose attenuation is a power of the column density, per-signal phase shifts.
"""
import numpy as np

from . import synthetic as _syn
from .likelihood import Likelihood as _Likelihood
from .pipeline import BatchedLikelihood as _BatchedLikelihood

_REGION_PARAMS = ("phase_shift", "super_colatitude", "super_radius", "super_temperature",
                  "omit_colatitude", "omit_radius", "omit_azimuth",
                  "cede_colatitude", "cede_radius", "cede_azimuth", "cede_temperature")


def _same(values, what):
    v0 = values[0]
    for v in values[1:]:
        same = np.array_equal(np.asarray(v), np.asarray(v0)) if isinstance(v0, np.ndarray) else v == v0
        if not same:
            raise NotImplementedError("xpsi_b200.from_xpsi: the hot regions differ in %s; one batched pipeline "
                                      "needs the same value for all of them" % what)
    return v0


def _resolve(obj, name):
    """The ``Parameter`` object ``obj[name]`` evaluates, or ``None`` when the object was constructed without it
    (read as 0: e.g. no ``cede_temperature`` when ``cede=False``, xpsi/HotRegion.py:300-480)."""
    try:
        return obj.get_param(name)
    except KeyError:
        return None


def _subspace_call(obj):
    """The reference's ``ParameterSubspace.__call__`` (xpsi/ParameterSubspace.py), bound to ``obj``: writing a
    parameter vector without triggering ``Likelihood.__call__`` -- what ``super(Likelihood, self).__call__(p)``
    does at xpsi/Likelihood.py:476."""
    for cls in type(obj).__mro__:
        if cls.__name__ == "ParameterSubspace":
            return lambda p: cls.__call__(obj, p)
    raise TypeError("the object handed to from_xpsi is not a ParameterSubspace (expected xpsi.Likelihood)")


def _attenuation_power_form(interstellar, mids):
    """``Interstellar.attenuation(energies)`` evaluated at two settings of its (single) parameter: the pipeline
    applies ``base ** power`` (xpsi/Interstellar.py:27-58 multiplies the signal by the user's ``attenuation``).
    Returns ``(parameter object or None, base)``; ``None`` = the attenuation does not depend on the vector."""
    params = list(getattr(interstellar, "_params", []))
    if len(params) != 1:
        raise NotImplementedError("xpsi_b200.from_xpsi: an Interstellar object with %d parameters (one expected: "
                                  "the column density)" % len(params))
    param = params[0]
    if getattr(param, "fixed", False):
        return None, np.array(interstellar.attenuation(mids), dtype=np.float64)
    old = getattr(param, "_value", None)
    try:
        param.value = 1.0
        base = np.array(interstellar.attenuation(mids), dtype=np.float64)
        param.value = 0.37
        probe = np.array(interstellar.attenuation(mids), dtype=np.float64)
    finally:
        if old is not None:
            param.value = old
    if not np.allclose(probe, base ** 0.37, rtol=1e-12, atol=0.0):
        raise NotImplementedError("xpsi_b200.from_xpsi: Interstellar.attenuation is not of the form "
                                  "base(E) ** column_density")
    return param, base


class _Model:
    """Static description of the reference model + the per-row parameter walk."""

    def __init__(self, likelihood, honour_signal_phase_shift):
        self.set_vector = _subspace_call(likelihood)
        self.names = list(likelihood.names)
        star = likelihood.star
        photospheres = star.photospheres
        if len(photospheres) != 1:
            raise NotImplementedError("xpsi_b200.from_xpsi: one photosphere per star is supported")
        self.ph = ph = photospheres[0]
        self.st = star.spacetime
        # Likelihood.signals hands back the sole Signal itself, or the list of per-photosphere lists
        # (xpsi/Likelihood.py:155-161)
        signals = likelihood.signals
        if not isinstance(signals, (list, tuple)):
            signals = [[signals]]
        elif not isinstance(signals[0], (list, tuple)):
            signals = [list(signals)]
        if len(signals) != 1:
            raise NotImplementedError("xpsi_b200.from_xpsi: one photosphere (one list of signals) is supported")
        self.signals = list(signals[0])
        self.honour_shift = bool(honour_signal_phase_shift)
        self.interstellar, self.column = None, None      # set while the signals are read (build_pipeline)
        hot, self.elsewhere, self.everywhere = ph.hot, ph.elsewhere, ph.everywhere
        self.regions = [] if hot is None else list(getattr(hot, "objects", [hot]))
        if self.everywhere is not None:
            if not getattr(self.everywhere, "_time_invariant", False):
                raise NotImplementedError("xpsi_b200.from_xpsi: Everywhere(time_invariant=False) is covered by the "
                                          "stand-alone general integrator, not by the batched pipeline")
        elif not self.regions:
            raise NotImplementedError("xpsi_b200.from_xpsi: the photosphere has neither hot regions nor Everywhere")
        for h in self.regions:
            if not getattr(h, "_symmetry", True):
                raise NotImplementedError("xpsi_b200.from_xpsi: HotRegion(symmetry=False) is covered by the "
                                          "stand-alone general integrator, not by the batched pipeline")
            if getattr(h, "_split", False):
                raise NotImplementedError("xpsi_b200.from_xpsi: split (Num5D) hot regions")
            if h.beam_opt != 0:
                raise NotImplementedError("xpsi_b200.from_xpsi: beaming parameters live in a user subclass of "
                                          "HotRegion; pass a fill function with extra_params instead")
        # the parameter objects behind every value the walk reads, resolved once (obj[name] = get_param(name).evaluate(obj)
        # with two linear scans over the names per call, xpsi/ParameterSubspace.py:112-130,185-189)
        self._p_st = [_resolve(self.st, n) for n in ("mass", "radius", "distance", "cos_inclination", "frequency")]
        self._p_mode = _resolve(ph, "mode_frequency")
        self._p_regions = [[_resolve(h, n) for n in _REGION_PARAMS] for h in self.regions]
        self._p_else = (_resolve(self.elsewhere, "elsewhere_temperature") if self.elsewhere is not None else
                        _resolve(self.everywhere, "temperature") if self.everywhere is not None else None)
        self._p_sig = [_resolve(s_, "phase_shift") for s_ in self.signals]
        # members: superseding member of every region, followed by its ceding member when the region has one
        self.member_region, self.member_is_cede = [], []
        for r, h in enumerate(self.regions):
            self.member_region.append(r); self.member_is_cede.append(0)
            if h._cede:
                self.member_region.append(r); self.member_is_cede.append(1)

    # ---- pipeline construction ---------------------------------------------------------------------------
    def _signal_kwargs(self, sig):
        inst, data = sig.instrument, sig.data
        a, b = sig._input_interval_range
        o0, o1 = sig._instrument_index_range_channels if hasattr(sig, "_instrument_index_range_channels") \
            else data.index_range
        matrix = np.ascontiguousarray(inst.construct_matrix()[o0:o1, a:b], dtype=np.float64)
        kw = dict(response=matrix, energy_edges=np.asarray(sig._energy_edges, dtype=np.float64),
                  counts=np.asarray(data.counts, dtype=np.float64), data_phases=np.asarray(data.phases, dtype=np.float64),
                  exposure_time=float(data.exposure_time), support=getattr(sig, "_support", None),
                  epsilon=float(getattr(sig, "_epsilon", 1.0e-3)), sigmas=float(getattr(sig, "_sigmas", 10.0)))
        att = None
        if sig.interstellar is not None:
            mids = 0.5 * (kw["energy_edges"][:-1] + kw["energy_edges"][1:])
            param, att = _attenuation_power_form(sig.interstellar, mids)
            if param is not None:
                if self.interstellar is None:
                    self.interstellar, self.column = sig.interstellar, param
                if sig.interstellar is not self.interstellar:
                    raise NotImplementedError("xpsi_b200.from_xpsi: the signals must share one Interstellar object")
        return kw, att

    def build_pipeline(self, max_batch, llzero, slim):
        sig0 = self.signals[0]
        energies = np.asarray(sig0.energies, dtype=np.float64)
        for s in self.signals[1:]:
            if not np.array_equal(np.asarray(s.energies), energies):
                raise NotImplementedError("xpsi_b200.from_xpsi: the signals of a photosphere share one energy array "
                                          "(xpsi/Likelihood.py:102-107)")
        self.interstellar = None
        kw0, att0 = self._signal_kwargs(sig0)
        common = dict(energies=energies, max_batch=max_batch, llzero=llzero, slim=slim, **kw0)
        if self.everywhere is not None:
            ev = self.everywhere
            pipe = _BatchedLikelihood(**common)
            atm = self.ph.everywhere_atmosphere if ev.atm_ext == 2 else None
            pipe.set_extras(everywhere=dict(sqrt_num_cells=ev.sqrt_num_cells, num_rays=ev.num_rays, atm_ext=ev.atm_ext,
                                            atmosphere=atm, image_order_limit=ev.image_order_limit), attenuation=att0)
        else:
            hs = self.regions
            leaves = _same([np.asarray(h.leaves) for h in hs], "leaves")
            phases = _same([np.asarray(h.phases) for h in hs], "phases")
            atm_ext = _same([h.atm_ext for h in hs], "atm_ext")
            if atm_ext not in (1, 2):
                raise NotImplementedError("xpsi_b200.from_xpsi: atmosphere extension %r (blackbody and Num4D are "
                                          "covered)" % (atm_ext,))
            pad = max(h._max_sqrt_num_cells for h in hs)
            if pad > 128:
                raise NotImplementedError("xpsi_b200.from_xpsi: max_sqrt_num_cells > 128")
            pipe = _BatchedLikelihood(member_component=self.member_region, max_rings=pad, max_azi=pad,
                                      n_rays=_same([h.num_rays for h in hs], "num_rays"), leaves=leaves, phases=phases,
                                      hot_atm_ext=atm_ext, hot_atmosphere=self.ph.hot_atmosphere if atm_ext == 2 else None,
                                      image_order_limit=_same([h.image_order_limit for h in hs], "image_order_limit"),
                                      **common)
            els = None
            if self.elsewhere is not None:
                e = self.elsewhere
                els = dict(sqrt_num_cells=e.sqrt_num_cells, num_rays=e.num_rays, atm_ext=e.atm_ext,
                           atmosphere=self.ph.elsewhere_atmosphere if e.atm_ext == 2 else None,
                           image_order_limit=e.image_order_limit)
            if els is not None or att0 is not None:
                pipe.set_extras(elsewhere=els, attenuation=att0)
        for s in self.signals[1:]:
            kw, att = self._signal_kwargs(s)
            pipe.add_signal(llzero=llzero, slim=slim, attenuation=att, **kw)
        return pipe

    # ---- parameter walk ----------------------------------------------------------------------------------
    def fill(self, pipe, P):
        P = np.atleast_2d(np.asarray(P, dtype=np.float64))
        B, R = P.shape[0], len(self.regions)
        st_v = np.empty((B, 5))
        reg_v = np.empty((B, max(R, 1), len(_REGION_PARAMS)))
        else_T = np.empty(B) if (self.elsewhere is not None or self.everywhere is not None) else None
        nh = np.empty(B) if self.interstellar is not None else None
        S = len(self.signals)
        sig_shift = np.zeros((B, S)) if (self.honour_shift and S >= 1) else None
        mode_f = np.empty(B)
        st, ph = self.st, self.ph
        for b in range(B):
            self.set_vector(P[b])
            st_v[b] = [p.evaluate(st) for p in self._p_st]
            mode_f[b] = self._p_mode.evaluate(ph)
            for r, h in enumerate(self.regions):
                reg_v[b, r] = [p.evaluate(h) if p is not None else 0.0 for p in self._p_regions[r]]
            if else_T is not None:
                else_T[b] = self._p_else.evaluate(self.elsewhere if self.elsewhere is not None else self.everywhere)
            if nh is not None:
                nh[b] = self.column.evaluate(self.interstellar)
            if sig_shift is not None:
                sig_shift[b] = [p.evaluate(s_) for p, s_ in zip(self._p_sig, self.signals)]
        if np.any(mode_f != mode_f[0]):
            raise NotImplementedError("xpsi_b200.from_xpsi: mode_frequency must be the same for a whole batch")
        M = max(len(self.member_region), 1)
        spots = pipe.new_spot_batch(B, float(mode_f[0]), num_cells=1, min_sqrt_num_cells=1, max_sqrt_num_cells=1)
        spots.set_spacetime(st_v[:, 0], st_v[:, 1], st_v[:, 2], st_v[:, 3], st_v[:, 4])
        if self.regions:
            cells = np.zeros((M, 3), dtype=np.int32)
            m = 0
            for r, h in enumerate(self.regions):
                v = {n: reg_v[:, r, i] for i, n in enumerate(_REGION_PARAMS)}
                cede = m + 1 if h._cede else None
                spots.set_region(m, cede, super_colatitude=v["super_colatitude"], super_radius=v["super_radius"],
                                 super_temperature=v["super_temperature"], omit_colatitude=v["omit_colatitude"],
                                 omit_radius=v["omit_radius"], omit_azimuth=v["omit_azimuth"],
                                 cede_colatitude=v["cede_colatitude"], cede_radius=v["cede_radius"],
                                 cede_azimuth=v["cede_azimuth"], cede_temperature=v["cede_temperature"],
                                 is_antiphased=bool(h._is_antiphased))
                spots.phase_shifts[:, r] = v["phase_shift"]
                for k in range(2 if h._cede else 1):
                    cells[m + k] = (h._num_cells, h._min_sqrt_num_cells, h._max_sqrt_num_cells)
                m += 2 if h._cede else 1
            spots.member_cells = cells
        extras = {}
        if else_T is not None:
            extras["else_temperature"] = else_T
        if nh is not None:
            extras["att_power"] = nh
        if sig_shift is not None and (S > 1 or np.any(sig_shift != 0.0)):
            extras["signal_shifts"] = sig_shift
        return spots, extras


def from_xpsi(likelihood, max_batch=256, slim=20.0, honour_signal_phase_shift=True, prior="inherit"):
    """GPU likelihood for the model held by a reference ``xpsi.Likelihood`` object.

    :param likelihood: the constructed ``xpsi.Likelihood`` (it is only read: its parameter values are overwritten
        by every evaluation, as they are by its own ``__call__``).
    :param max_batch: parameter vectors per launch sequence.
    :param slim: the ``slim`` argument of ``eval_marginal_likelihood`` the model's ``Signal`` subclass passes
        (default_background_marginalisation.pyx:450-470; the example ``CustomSignal`` classes leave it at 20).
    :param honour_signal_phase_shift: add each signal's own ``phase_shift`` parameter to the hot regions' shifts
        (``Signal.shifts``, xpsi/Signal.py:581-583).  The example ``CustomSignal.__call__`` passes ``self._shifts``
        and therefore ignores it; with a parameter fixed at zero (the default) both conventions agree.
    :param prior: ``"inherit"`` takes ``likelihood.prior`` when the object has one; ``None`` for no prior.
    """
    model = _Model(likelihood, honour_signal_phase_shift)
    llzero = float(getattr(likelihood, "llzero", -1.0e90))
    pipe = model.build_pipeline(int(max_batch), llzero, float(slim))
    if prior == "inherit":
        try:
            prior = likelihood.prior
        except AttributeError:
            prior = None
    out = _Likelihood(pipe, model.fill, prior=prior, llzero=llzero, max_batch=max_batch)
    out.names = model.names
    out.model = model
    return out
