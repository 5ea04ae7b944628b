"""ctypes binding of libxpsi_b200.so (the C ABI in include/xpsi_b200.h).

There is no CPU fallback: importing this module fails loudly when the shared
library has not been built, and every call fails when no CUDA device is usable.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# XPSI_B200_LIB: development only (timing kernel variants built by dev/build_variants.sh side by side)
LIB_PATH = os.environ.get("XPSI_B200_LIB") or os.path.join(_HERE, "libxpsi_b200.so")

if not os.path.isfile(LIB_PATH):
    raise ImportError(
        "xpsi_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C xpsi_b200/csrc`; the hot path has no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)

OK = 0
ENUMERICAL = 1
EUNSUPPORTED = 3
ESLIM = 11
EQUADRATURE = 12

INTERPOLANTS = {'Akima': 0, 'Steffen': 1, 'Cubic': 2}


class XpsiB200Error(RuntimeError):
    """API / CUDA failure inside libxpsi_b200 (negative return codes)."""


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


def as_f8(a, ndim=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if ndim is not None and a.ndim != ndim:
        raise ValueError("expected a %d-D array, got shape %r" % (ndim, a.shape))
    return a


def as_i4(a, ndim=None):
    a = np.ascontiguousarray(a, dtype=np.int32)
    if ndim is not None and a.ndim != ndim:
        raise ValueError("expected a %d-D array, got shape %r" % (ndim, a.shape))
    return a


class PipelineConfig(C.Structure):
    _fields_ = [
        ("n_components", C.c_int), ("n_members", C.c_int), ("member_component", c_int_p),
        ("max_rings", C.c_int), ("max_azi", C.c_int), ("n_rays", C.c_int), ("n_params", C.c_int),
        ("n_energies", C.c_int), ("energies", c_double_p),
        ("n_leaves", C.c_int), ("leaves", c_double_p),
        ("n_phases", C.c_int), ("phases", c_double_p),
        ("hot_atm_ext", C.c_int), ("hot_atmosphere", C.c_void_p),
        ("image_order_limit", C.c_int), ("phase_interpolant", C.c_int),
        ("n_in", C.c_int), ("energy_edges", c_double_p),
        ("n_chan", C.c_int), ("response", c_double_p),
        ("n_bins", C.c_int), ("data_phases", c_double_p),
        ("counts", c_double_p), ("support", c_double_p),
        ("exposure_time", C.c_double), ("epsilon", C.c_double), ("sigmas", C.c_double),
        ("llzero", C.c_double), ("slim", C.c_double), ("allow_negative", C.c_int),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("omega", c_double_p), ("inclination", c_double_p), ("d_sq", c_double_p),
        ("phase_shifts", c_double_p),
        ("n_rings", c_int_p), ("n_azi", c_int_p),
        ("cellArea", c_double_p), ("phi", c_double_p), ("theta", c_double_p),
        ("radial", c_double_p), ("r_s_over_r", c_double_p), ("srcParams", c_double_p),
        ("deflection", c_double_p), ("cos_alpha", c_double_p), ("lag", c_double_p),
        ("maxDeflection", c_double_p), ("cos_gamma", c_double_p),
    ]


class SpotBatch(C.Structure):
    _fields_ = [
        ("R_eq", c_double_p), ("r_s", c_double_p), ("epsilon", c_double_p), ("zeta", c_double_p),
        ("omega", c_double_p), ("inclination", c_double_p), ("d_sq", c_double_p), ("phase_shifts", c_double_p),
        ("colatitude", c_double_p), ("ang_radius", c_double_p), ("temperature", c_double_p),
        ("phi_shift", c_double_p), ("mode_frequency", C.c_double),
        ("num_cells", C.c_int), ("min_sqrt_num_cells", C.c_int), ("max_sqrt_num_cells", C.c_int),
        ("hole_radius", c_double_p), ("hole_colatitude", c_double_p), ("hole_azimuth", c_double_p),
        ("partner", c_int_p), ("is_cede", c_int_p), ("extra_params", c_double_p),
        ("member_cells", c_int_p),
    ]


class PipelineExtras(C.Structure):
    _fields_ = [
        ("elsewhere", C.c_int), ("else_sqrt_num_cells", C.c_int), ("else_num_rays", C.c_int),
        ("else_atm_ext", C.c_int), ("else_image_order_limit", C.c_int), ("elsewhere_atmosphere", C.c_void_p),
        ("attenuation", c_double_p), ("beam_opt", C.c_int),
    ]


class BatchExtras(C.Structure):
    _fields_ = [(f, c_double_p) for f in (
        "att_power", "else_temperature", "else_cellArea", "else_radial", "else_r_s_over_r", "else_theta",
        "else_phi", "else_srcParams", "else_deflection", "else_cos_alpha", "else_maxDeflection", "else_cos_gamma",
        "correction_srcParams")]


class SignalConfig(C.Structure):
    _fields_ = [
        ("n_in", C.c_int), ("energy_edges", c_double_p),
        ("n_chan", C.c_int), ("response", c_double_p),
        ("n_bins", C.c_int), ("data_phases", c_double_p),
        ("counts", c_double_p), ("support", c_double_p),
        ("exposure_time", C.c_double), ("epsilon", C.c_double), ("sigmas", C.c_double),
        ("llzero", C.c_double), ("slim", C.c_double), ("allow_negative", C.c_int),
        ("attenuation", c_double_p),
    ]


def _proto(name, restype, argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = argtypes
    return f


_proto("xpsi_b200_last_error", C.c_char_p, [])
_proto("xpsi_b200_device_count", C.c_int, [])
_proto("xpsi_b200_set_device", C.c_int, [C.c_int])
_proto("xpsi_b200_counters", None, [C.POINTER(C.c_longlong)] * 3)
_proto("xpsi_b200_stream", C.c_void_p, [])
_proto("xpsi_b200_atmosphere_create", C.c_void_p,
       [c_double_p, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, c_double_p])
_proto("xpsi_b200_atmosphere_destroy", None, [C.c_void_p])
_proto("xpsi_b200_integrate_azimuthal_invariance", C.c_int,
       [C.c_double] * 4 + [C.c_int, C.c_int] + [c_double_p] * 5 + [c_double_p, C.c_int, c_int_p, c_double_p,
        C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
        C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, c_double_p,
        C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, c_double_p])
_proto("xpsi_b200_integrate_general", C.c_int,
       [C.c_double] * 4 + [C.c_int, C.c_int] + [c_double_p] * 5 + [c_double_p, C.c_int, c_int_p, c_double_p,
        C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
        C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, c_double_p,
        C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, c_double_p])
_proto("xpsi_b200_pipeline_set_extras", C.c_int, [C.c_void_p, C.POINTER(PipelineExtras)])
_proto("xpsi_b200_pipeline_upload_extras", C.c_int, [C.c_void_p, C.c_int, C.POINTER(BatchExtras)])
_proto("xpsi_b200_pipeline_fetch_elsewhere", C.c_int, [C.c_void_p, C.c_int, c_double_p])
_proto("xpsi_b200_intensity", C.c_int,
       [C.c_int, c_double_p, c_double_p, c_double_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p])
_proto("xpsi_b200_integrate_time_invariance", C.c_int,
       [C.c_double] * 4 + [C.c_int, C.c_double] + [c_double_p] * 5 + [C.c_int, C.c_int] + [c_double_p] * 4 +
       [C.c_int, c_double_p, C.c_void_p, C.c_int, C.c_int, c_double_p])
_proto("xpsi_b200_energy_integrator", C.c_int,
       [c_double_p, C.c_int, C.c_int, c_double_p, c_double_p, C.c_int, C.c_int, c_double_p])
_proto("xpsi_b200_phase_integrator", C.c_int,
       [C.c_double, c_double_p, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, C.c_double, C.c_int, C.c_int,
        c_double_p])
_proto("xpsi_b200_phase_interpolator", C.c_int,
       [c_double_p, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, C.c_double, C.c_int, C.c_int, c_double_p])
_proto("xpsi_b200_energy_interpolator", C.c_int,
       [c_double_p, C.c_int, C.c_int, c_double_p, c_double_p, C.c_int, C.c_int, c_double_p])
_proto("xpsi_b200_interstellar_attenuate", C.c_int, [c_double_p, C.c_int, C.c_int, c_double_p])
_proto("xpsi_b200_instrument_fold", C.c_int,
       [c_double_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, C.c_int, c_double_p])
_proto("xpsi_b200_precomputation", C.c_int, [c_int_p, C.c_int, C.c_int, c_double_p])
_proto("xpsi_b200_eval_marginal_likelihood", C.c_int,
       [C.c_double, c_double_p, C.c_int, c_double_p, C.c_int, C.POINTER(c_double_p), C.c_int, C.POINTER(c_double_p),
        c_int_p, c_double_p, c_double_p, c_double_p, C.c_double, C.c_double, C.c_double, c_int_p,
        C.c_double, c_double_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p])
_proto("xpsi_b200_poisson_likelihood_given_background", C.c_int,
       [C.c_double, c_double_p, C.c_int, c_double_p, C.c_int, C.POINTER(c_double_p), C.c_int, C.POINTER(c_double_p),
        c_int_p, c_double_p, c_double_p, c_double_p, c_int_p, C.c_int, c_double_p, c_double_p])
_proto("xpsi_b200_pipeline_create", C.c_void_p, [C.POINTER(PipelineConfig), C.c_int])
_proto("xpsi_b200_pipeline_destroy", None, [C.c_void_p])
_proto("xpsi_b200_pipeline_eval", C.c_int, [C.c_void_p, C.c_int, C.POINTER(Batch), c_double_p, c_int_p])
_proto("xpsi_b200_pipeline_upload", C.c_int, [C.c_void_p, C.c_int, C.POINTER(Batch)])
_proto("xpsi_b200_pipeline_eval_resident", C.c_int, [C.c_void_p, C.c_int])
_proto("xpsi_b200_pipeline_download", C.c_int, [C.c_void_p, C.c_int, c_double_p, c_int_p])
_proto("xpsi_b200_pipeline_fetch", C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p])
_proto("xpsi_b200_fp64_peak_tflops", C.c_int, [c_double_p])
_proto("xpsi_b200_pipeline_work_counters", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_ulonglong)])
_proto("xpsi_b200_pipeline_embed_spots", C.c_int, [C.c_void_p, C.c_int, C.POINTER(SpotBatch)])
_proto("xpsi_b200_pipeline_eval_spots_resident", C.c_int, [C.c_void_p, C.c_int])
_proto("xpsi_b200_pipeline_eval_spots", C.c_int, [C.c_void_p, C.c_int, C.POINTER(SpotBatch), c_double_p, c_int_p])
_proto("xpsi_b200_pipeline_fetch_embed", C.c_int, [C.c_void_p, C.c_int, c_int_p] + [c_double_p] * 10)
_proto("xpsi_b200_pipeline_stage_ms", C.c_int, [C.c_void_p, C.POINTER(C.c_float)])
_proto("xpsi_b200_pipeline_set_deterministic", C.c_int, [C.c_void_p, C.c_int])
_proto("xpsi_b200_pipeline_sweep_upload", C.c_int, [C.c_void_p, C.c_longlong, C.POINTER(SpotBatch), c_double_p, c_double_p])
_proto("xpsi_b200_pipeline_sweep_run", C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong])
_proto("xpsi_b200_pipeline_sweep_download", C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, c_double_p, c_int_p])
_proto("xpsi_b200_pipeline_sweep_results", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)])
_proto("xpsi_b200_pipeline_add_signal", C.c_int, [C.c_void_p, C.POINTER(SignalConfig)])
_proto("xpsi_b200_pipeline_n_signals", C.c_int, [C.c_void_p])
_proto("xpsi_b200_pipeline_upload_signal_shifts", C.c_int, [C.c_void_p, C.c_int, c_double_p])
_proto("xpsi_b200_pipeline_sweep_upload_signal_shifts", C.c_int, [C.c_void_p, C.c_longlong, c_double_p])
_proto("xpsi_b200_pipeline_fetch_signal", C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p])

EXPORTED = [
    "xpsi_b200_last_error", "xpsi_b200_device_count", "xpsi_b200_set_device", "xpsi_b200_counters",
    "xpsi_b200_stream", "xpsi_b200_atmosphere_create", "xpsi_b200_atmosphere_destroy",
    "xpsi_b200_integrate_azimuthal_invariance", "xpsi_b200_integrate_general", "xpsi_b200_intensity",
    "xpsi_b200_pipeline_set_extras", "xpsi_b200_pipeline_upload_extras", "xpsi_b200_pipeline_fetch_elsewhere", "xpsi_b200_energy_integrator",
    "xpsi_b200_instrument_fold", "xpsi_b200_precomputation", "xpsi_b200_eval_marginal_likelihood",
    "xpsi_b200_pipeline_create", "xpsi_b200_pipeline_destroy", "xpsi_b200_pipeline_eval",
    "xpsi_b200_pipeline_upload", "xpsi_b200_pipeline_eval_resident", "xpsi_b200_pipeline_download",
    "xpsi_b200_pipeline_fetch", "xpsi_b200_pipeline_stage_ms", "xpsi_b200_fp64_peak_tflops",
    "xpsi_b200_pipeline_work_counters", "xpsi_b200_phase_integrator", "xpsi_b200_phase_interpolator",
    "xpsi_b200_energy_interpolator", "xpsi_b200_integrate_time_invariance", "xpsi_b200_interstellar_attenuate",
    "xpsi_b200_pipeline_embed_spots", "xpsi_b200_pipeline_eval_spots", "xpsi_b200_pipeline_fetch_embed",
    "xpsi_b200_pipeline_eval_spots_resident", "xpsi_b200_poisson_likelihood_given_background",
    "xpsi_b200_pipeline_sweep_upload", "xpsi_b200_pipeline_sweep_run", "xpsi_b200_pipeline_sweep_download",
    "xpsi_b200_pipeline_sweep_results", "xpsi_b200_pipeline_set_deterministic",
    "xpsi_b200_pipeline_add_signal", "xpsi_b200_pipeline_n_signals", "xpsi_b200_pipeline_upload_signal_shifts",
    "xpsi_b200_pipeline_sweep_upload_signal_shifts", "xpsi_b200_pipeline_fetch_signal",
]


def last_error():
    return lib.xpsi_b200_last_error().decode("utf-8", "replace")


class XpsiB200NumericalError(XpsiB200Error):
    """The reference's numerical ``ERROR`` return reached a wrapper that has no ``(1, None)`` convention."""


def check(rc, allow=()):
    """Raise on every non-zero return code that the caller does not list in ``allow``.

    Negative codes are API / CUDA failures (``XpsiB200Error``); ``EUNSUPPORTED`` is a configuration outside
    the kernels' coverage (``NotImplementedError``, never a silent wrong answer); the remaining positive
    codes (``ENUMERICAL``, ``ESLIM``, ``EQUADRATURE``) are numerical outcomes that a wrapper must either map
    onto the reference's convention itself (and then allow here) or surface as an exception.
    """
    if rc == OK or rc in allow:
        return rc
    if rc < 0:
        raise XpsiB200Error("libxpsi_b200 error %d: %s" % (rc, last_error()))
    if rc == EUNSUPPORTED:
        raise NotImplementedError("xpsi_b200: " + last_error())
    raise XpsiB200NumericalError("libxpsi_b200 numerical status %d: %s" % (rc, last_error()))


def counters():
    k, h, d = C.c_longlong(), C.c_longlong(), C.c_longlong()
    lib.xpsi_b200_counters(C.byref(k), C.byref(h), C.byref(d))
    return k.value, h.value, d.value


class Atmosphere:
    """Device-resident preloaded atmosphere (the reference's ``_preloaded``).

    Built from the reference's tuple ``(logT, logg, mu, logE, buf)``
    (xpsi/Photosphere.py:208-217); cached per tuple identity by ``get``.
    """
    _cache = {}

    def __init__(self, table):
        logT, logg, mu, logE, buf = [as_f8(t, 1) for t in table]
        if buf.size != logT.size * logg.size * mu.size * logE.size:
            raise ValueError("atmosphere buffer size does not match its axes")
        self._keep = (logT, logg, mu, logE, buf)
        self.handle = lib.xpsi_b200_atmosphere_create(dptr(logT), logT.size, dptr(logg), logg.size,
                                                      dptr(mu), mu.size, dptr(logE), logE.size, dptr(buf))
        if not self.handle:
            raise XpsiB200Error("atmosphere_create failed: %s" % last_error())

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib.xpsi_b200_atmosphere_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @classmethod
    def get(cls, table):
        if not table:
            return None
        if isinstance(table, Atmosphere):
            return table
        key = tuple(id(t) for t in table)
        hit = cls._cache.get(key)
        if hit is None or hit[1] is not table[-1]:
            hit = (cls(table), table[-1])
            cls._cache[key] = hit
        return hit[0]
