"""Recipe for oracle/_ref: the UNMODIFIED reference modules of the hot path, made importable on a box
that has neither /root/reference nor matplotlib.

    python oracle/build_ref.py            # in the build container (needs /root/reference)

copies libs/{__init__,FingerprintLib,OTlib,ricker_util,ricker_util_opt,myGP,loc_cmt_util,loc_cmt_util_opt}.py byte
for byte into the
git-ignored oracle/_ref/libs/ (nothing is edited; `cmp` against the originals is part of the recipe) and
writes inert stub packages for the plotting imports those modules make at module scope (matplotlib, pylab,
mpl_toolkits: libs/FingerprintLib.py:15-18, libs/OTlib.py:18, libs/ricker_util.py:11-13).  oracle/_ref/ is
listed in .gitignore (reference sources never enter the history) but not in .gpurunignore, so it travels to
the GPU box with the snapshot like the built libwfot.so.

Users: bench.py (`--impl reference` and the `cpu_baseline` leg: kind "reference") and
tests/test_gpu_dropin.py (the unmodified libs.ricker_util.optfunc and libs.loc_cmt_util.optfunc_OT running over the
B200 shim; loc_cmt_util imports the third-party pyprop8, absent from the image: oracle/pyprop8_stub.py).  Like the rest
of oracle/ it is test / measurement infrastructure: the product never imports it.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")
MODULES = ["__init__.py", "FingerprintLib.py", "OTlib.py", "ricker_util.py", "ricker_util_opt.py", "myGP.py",
           "loc_cmt_util.py", "loc_cmt_util_opt.py"]

# notebooks whose code cells tests/test_gpu_dropin.py executes over the shim (copied unmodified, like the modules)
NOTEBOOKS = ["Point_mass_demo_Fig_5.ipynb", "Ricker_waveform_derivatives.ipynb", "Ricker_Figs_3_8.ipynb"]

STUB = '''"""Inert stand-in for a plotting package the reference imports at module scope (written by
oracle/build_ref.py; never used by the hot path)."""
import sys as _sys
import types as _types


class _Placeholder:
    def __call__(self, *a, **k):
        return _Placeholder()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder()

    def __iter__(self):
        return iter(())


class _Anything(_types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder()


def _install(name):
    m = _Anything(name)
    m.__path__ = []
    _sys.modules[name] = m
    return m
'''


def available():
    return os.path.exists(os.path.join(DST, "libs", "OTlib.py"))


def build(verbose=True):
    src = os.path.join(REF, "libs")
    if not os.path.isdir(src):
        if verbose:
            print("oracle/_ref: %s not present (GPU box): using the prebuilt copy" % src if available()
                  else "oracle/_ref: reference not present and no prebuilt copy")
        return available()
    os.makedirs(os.path.join(DST, "libs"), exist_ok=True)
    for m in MODULES:
        shutil.copyfile(os.path.join(src, m), os.path.join(DST, "libs", m))
        assert filecmp.cmp(os.path.join(src, m), os.path.join(DST, "libs", m), shallow=False), m
    os.makedirs(os.path.join(DST, "notebooks"), exist_ok=True)
    for nb in NOTEBOOKS:
        shutil.copyfile(os.path.join(REF, nb), os.path.join(DST, "notebooks", nb))
        assert filecmp.cmp(os.path.join(REF, nb), os.path.join(DST, "notebooks", nb), shallow=False), nb
    with open(os.path.join(DST, "_plot_stubs.py"), "w") as f:
        f.write(STUB)
    if verbose:
        print("oracle/_ref: %d reference modules copied unmodified from %s" % (len(MODULES), src))
    return True


def import_reference():
    """-> (FingerprintLib, OTlib, ricker_util) of the unmodified reference in oracle/_ref (None if absent).
    Plotting packages that are missing from the image are replaced by inert stubs first; sklearn and scipy,
    which the reference also imports at module scope, are present in the image."""
    if not available():
        return None
    import importlib
    import warnings
    warnings.filterwarnings("ignore")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    stubs = importlib.import_module("_plot_stubs")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "pylab",
                 "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                stubs._install(name)
    for k in [k for k in sys.modules if k == "libs" or k.startswith("libs.")]:
        del sys.modules[k]                       # a shim installed under the same prefix must not leak in
    fp = importlib.import_module("libs.FingerprintLib")
    OT = importlib.import_module("libs.OTlib")
    ru = importlib.import_module("libs.ricker_util")
    return fp, OT, ru


def import_cmt():
    """-> the unmodified libs.loc_cmt_util of oracle/_ref, bound to whatever libs.FingerprintLib / libs.OTlib are
    installed in sys.modules at this point (the reference's own after import_reference(), the B200 shim after
    adapters.install("libs")).  pyprop8 (libs/loc_cmt_util.py:9,12) is replaced by oracle/pyprop8_stub.py if absent."""
    import importlib
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import pyprop8_stub
    pyprop8_stub.install()
    for k in ("libs.loc_cmt_util", "libs.loc_cmt_util_opt"):
        sys.modules.pop(k, None)
    return importlib.import_module("libs.loc_cmt_util")


if __name__ == "__main__":
    ok = build()
    sys.exit(0 if ok else 1)
