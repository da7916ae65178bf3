"""Import the UNMODIFIED reference (`/root/reference/libs`) in the build container.

Only the fixture generator (`make_golden.py`) and ad-hoc validation scripts use
this.  It never runs on the GPU box (`/root/reference` does not exist there);
nothing under tests/ imports it at test time.

matplotlib / pylab / mpl_toolkits are absent from this image, and the reference
imports them at module scope (libs/OTlib.py:18, libs/FingerprintLib.py:15-18,
libs/ricker_util.py:11-13), so permissive stub modules are injected first.
"""
import sys
import types
import warnings

REFERENCE_ROOT = "/root/reference"


class _Anything(types.ModuleType):
    """Module stub: any attribute is a callable/inert placeholder."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder()


class _Placeholder:
    def __call__(self, *a, **k):
        return _Placeholder()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder()


def import_reference():
    warnings.filterwarnings("ignore")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm",
                 "matplotlib.colors", "pylab", "mpl_toolkits",
                 "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from libs import FingerprintLib as fp, OTlib as OT, ricker_util as ru
    return fp, OT, ru
