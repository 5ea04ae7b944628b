"""Import helper for the reference build in oracle/_ref (TEST INFRASTRUCTURE).

``xpsi`` imports astropy / matplotlib / h5py-style packages at module top
(xpsi/Signal.py:20, xpsi/Instrument.py:14-15, xpsi/Data.py:9,
xpsi/shell_interpolator.py:16-17) although the likelihood path never calls
them.  They are absent here, so inert stand-in modules are injected before
``import xpsi``.  Nothing under xpsi_b200/ may import this file.
"""
import importlib
import importlib.abc
import importlib.machinery
import io
import os
import sys
import types
import contextlib

REF_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

# samplers / post-processing packages are guarded by try/except ImportError in
# the reference and are left absent
_STUB_ROOTS = ("astropy", "matplotlib", "mpl_toolkits")


class _Anything:
    """Attribute sink: any attribute / call / subscript yields another sink."""
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Anything()
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()
    def __getitem__(self, k): return _Anything()
    def __iter__(self): return iter(())
    def __mro_entries__(self, bases): return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        # xpsi.pixelmesh (sky maps, gsl_odeiv2) is not built: SURVEY.md s2 row 25
        if fullname.split(".")[0] in _STUB_ROOTS or fullname.startswith("xpsi.pixelmesh"):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None
    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m
    def exec_module(self, module):
        pass


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "xpsi", "__init__.py")) and \
        os.path.isdir(os.path.join(REF_ROOT, "xpsi", "cellmesh"))


def import_reference(quiet=True):
    """Return the reference ``xpsi`` package built by oracle/build_ref.py."""
    if not available():
        raise ImportError("oracle/_ref is not built (run python oracle/build_ref.py)")
    if "xpsi" in sys.modules and getattr(sys.modules["xpsi"], "__file__", "").startswith(REF_ROOT):
        return sys.modules["xpsi"]
    for name in _STUB_ROOTS:
        try:
            importlib.import_module(name)      # a real install wins
        except Exception:
            pass
    # The reference predates NumPy 2 (xpsi/Signal.py:235 uses ``numpy.infty``);
    # restore the removed aliases in this process instead of editing sources.
    import numpy as _np
    for alias, target in (("infty", "inf"), ("Inf", "inf"), ("float_", "float64"),
                          ("product", "prod"), ("NaN", "nan")):
        if alias not in _np.__dict__:
            setattr(_np, alias, getattr(_np, target))
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    sink = io.StringIO()
    ctx = contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()
    with ctx:
        xpsi = importlib.import_module("xpsi")
    return xpsi
