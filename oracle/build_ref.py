#!/usr/bin/env python
"""Build the UNMODIFIED reference hot path into oracle/_ref/ (git-ignored).

TEST INFRASTRUCTURE ONLY -- nothing under xpsi_b200/ may import oracle/.

The reference (X-PSI v3.3.0, /root/reference) is Cython + GSL and GSL is not
installed here (setup.py:89-94 aborts without gsl-config).  This recipe
therefore side-steps the reference's own build system:

  1. cythonize each hot-path ``.pyx`` *where it lies* under /root/reference
     (sources are only read) into generated C under oracle/_ref/build/;
  2. compile the C against the from-scratch GSL-subset shim in oracle/gslshim/
     (same flags as setup.py:118-125 except -march=native -> x86-64-v3 so the
     binaries also run on the GPU box's host CPU);
  3. install the reference's pure-Python package files next to the extension
     modules (the equivalent of ``pip install --target``), so that
     ``xpsi.Likelihood`` can be driven as the CPU baseline and as the source of
     golden vectors (tests/golden/make_golden.py).

Only oracle/_ref/ is written.  pixelmesh (sky maps, needs gsl_odeiv2) is not
built; it is never on the likelihood path.

Usage: python oracle/build_ref.py [--reference /root/reference] [-j N]
"""
import argparse
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "gslshim")

# setup.py:207-255 minus pixelmesh
MODULES = [
    "xpsi.surface_radiation_field.effective_gravity_universal",
    "xpsi.cellmesh.mesh_tools",
    "xpsi.cellmesh.mesh",
    "xpsi.cellmesh.polar_mesh",
    "xpsi.cellmesh.global_mesh",
    "xpsi.cellmesh.rays",
    "xpsi.tools.energy_interpolator",
    "xpsi.tools.energy_integrator",
    "xpsi.tools.phase_integrator",
    "xpsi.tools.phase_interpolator",
    "xpsi.tools.synthesise",
    "xpsi.tools.compute_expected_counts",
    "xpsi.tools.core",
    "xpsi.likelihoods.default_background_marginalisation",
    "xpsi.likelihoods._poisson_likelihood_given_background",
    "xpsi.likelihoods._gaussian_likelihood_given_background_IQU",
    "xpsi.likelihoods._gaussian_likelihood_QnUn",
    "xpsi.surface_radiation_field.core",
    "xpsi.surface_radiation_field.preload",
    "xpsi.surface_radiation_field.hot_user",
    "xpsi.surface_radiation_field.hot_BB",
    "xpsi.surface_radiation_field.hot_Num4D",
    "xpsi.surface_radiation_field.hot_BB_burst",
    "xpsi.surface_radiation_field.hot_Num2D",
    "xpsi.surface_radiation_field.hot_Num2D_split",
    "xpsi.surface_radiation_field.hot_Num5D_split",
    "xpsi.surface_radiation_field.hot_wrapper",
    "xpsi.surface_radiation_field.elsewhere_user",
    "xpsi.surface_radiation_field.elsewhere_wrapper",
    "xpsi.surface_radiation_field.local_variables",
    "xpsi.cellmesh.integrator",
    "xpsi.cellmesh.common_functions",
    "xpsi.cellmesh.integratorIQU",
    "xpsi.cellmesh.integrator_for_azimuthal_invariance",
    "xpsi.cellmesh.integrator_for_azimuthal_invariance_split",
    "xpsi.cellmesh.integratorIQU_for_azimuthal_invariance",
    "xpsi.cellmesh.integratorIQU_for_azimuthal_invariance_split",
    "xpsi.cellmesh.integrator_for_time_invariance",
]

CFLAGS = ["-fopenmp", "-march=x86-64-v3", "-O3", "-funroll-loops", "-fPIC",
          "-Wno-unused-function", "-Wno-uninitialized", "-Wno-cpp",
          "-Wno-maybe-uninitialized", "-Wno-unused-variable"]


def run(cmd):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + p.stdout[-4000:] + "\n")
        raise SystemExit("build_ref: command failed")
    return p.stdout


def build_shim():
    os.makedirs(os.path.join(OUT, "lib"), exist_ok=True)
    lib = os.path.join(OUT, "lib", "libgslshim.so")
    run(["gcc", "-O2", "-fPIC", "-shared", "-I" + os.path.join(SHIM, "include"),
         "-I" + SHIM, os.path.join(SHIM, "gslshim.c"), "-o", lib, "-lm"])
    return lib


def build_module(ref, mod):
    import numpy
    rel = mod.replace(".", os.sep)
    pyx = os.path.join(ref, rel + ".pyx")
    c_out = os.path.join(OUT, "build", rel + ".c")
    os.makedirs(os.path.dirname(c_out), exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so_out = os.path.join(OUT, rel + ext)
    os.makedirs(os.path.dirname(so_out), exist_ok=True)
    if os.path.exists(so_out) and os.path.getmtime(so_out) > os.path.getmtime(pyx):
        return mod, "cached"
    run([sys.executable, "-m", "cython", "-3", "--module-name", mod,
         "-I", ref, "-I", os.path.join(ref, "xpsi", "include"),
         "-o", c_out, pyx])
    run(["gcc", "-shared"] + CFLAGS +
        ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
         "-I" + os.path.join(SHIM, "include"),
         "-I" + os.path.join(ref, "xpsi", "include"),
         c_out, "-o", so_out,
         "-L" + os.path.join(OUT, "lib"), "-lgslshim", "-lm",
         "-Wl,-rpath,$ORIGIN/../../lib"])
    return mod, "built"


def install_python(ref):
    """Install the pure-Python half of the package (pip --target equivalent)."""
    src = os.path.join(ref, "xpsi")
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "pixelmesh", "include")]
        for f in files:
            if f.endswith(".py"):
                dst = os.path.join(OUT, os.path.relpath(root, ref), f)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), dst)
    # the examples_fast custom model classes + data drive the golden run
    ex = os.path.join(ref, "examples", "examples_fast")
    for sub in ("Modules", "Data"):
        d = os.path.join(OUT, "examples_fast", sub)
        os.makedirs(d, exist_ok=True)
        for f in os.listdir(os.path.join(ex, sub)):
            if f.endswith((".py", ".dat")):
                shutil.copyfile(os.path.join(ex, sub, f), os.path.join(d, f))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("-j", type=int, default=len(os.sched_getaffinity(0)))
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(a.reference, "xpsi")):
        print("build_ref: %s not present; keeping any prebuilt oracle/_ref" % a.reference)
        return 0
    build_shim()
    with ThreadPoolExecutor(a.j) as ex:
        for mod, how in ex.map(lambda m: build_module(a.reference, m), MODULES):
            print("  %-70s %s" % (mod, how))
    install_python(a.reference)
    with open(os.path.join(OUT, "BUILD_INFO.txt"), "w") as fh:
        fh.write("reference: X-PSI 3.3.0 (unmodified .pyx/.py from %s)\n" % a.reference)
        fh.write("gsl: from-scratch subset shim oracle/gslshim (NOT GNU GSL)\n")
        fh.write("cflags: %s\n" % " ".join(CFLAGS))
    print("build_ref: done ->", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
