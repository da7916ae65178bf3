"""GPU: the UNMODIFIED reference adapters running over the B200 shim.

oracle/_ref holds byte-for-byte copies of the reference's libs/ricker_util.py (+ the modules it imports),
made by oracle/build_ref.py in the build container and shipped to the GPU box with the snapshot (git-ignored).
`adapters.install("libs")` substitutes waveform_ot_b200.FingerprintLib / OTlib for libs.FingerprintLib /
libs.OTlib; libs.ricker_util then runs unchanged: rickerwavelet -> BuildOTobjfromWaveform ->
CalcWasserWaveform -> chain rule (libs/ricker_util.py:373-404).  Expected values: tests/golden/
ricker_forward.npz (F, G), produced by the same calls on the unmodified reference end to end."""
import importlib
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def ref_over_shim():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    sys.path.insert(0, ROOT)
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py in the build container)")
    fp_ref, OT_ref, ru_ref = build_ref.import_reference()      # also installs plotting stubs, sys.path
    saved = {k: v for k, v in sys.modules.items() if k == "libs" or k.startswith("libs.")}
    for k in saved:
        del sys.modules[k]
    importlib.import_module("libs")                            # the reference package itself (oracle/_ref/libs)
    from waveform_ot_b200 import adapters
    fpm, otm = adapters.install("libs")                        # libs.FingerprintLib / libs.OTlib -> B200 shim
    ru = importlib.import_module("libs.ricker_util")           # UNMODIFIED reference source
    assert ru.fp is fpm and ru.OT is otm
    assert os.path.realpath(ru.__file__).startswith(os.path.realpath(os.path.join(ROOT, "oracle", "_ref")))
    yield ru, otm
    for k in [k for k in sys.modules if k == "libs" or k.startswith("libs.")]:
        del sys.modules[k]
    sys.modules.update(saved)


def test_unmodified_ricker_util_optfunc_over_shim(ref_over_shim, golden):
    ru, _ = ref_over_shim
    g = golden("ricker_forward")
    grid = tuple(g["grid"][:4]) + (int(g["grid"][4]), int(g["grid"][5]))
    lam, alpha = float(g["lam"]), float(g["alpha"])
    wfo, tgt = ru.BuildOTobjfromWaveform(g["to"], g["wo"], grid, lambdav=lam)      # observed window (shim objects)
    ru.ricker_util_opt.init()
    for x, F, G in zip(g["X"], g["F"], g["G"]):
        data = [tgt, "W2", [-2.0, 2.0], grid, lam, False, alpha, 45.0]
        w2, deriv = ru.optfunc(x, data)                                            # libs/ricker_util.py:373-404
        assert w2 == pytest.approx(float(F), rel=1e-9)
        np.testing.assert_allclose(deriv, G, rtol=1e-7, atol=1e-10)
    assert len(ru.ricker_util_opt.Wdata) == len(g["X"])                           # the reference's history side effect


def test_unmodified_optfunc_raises_on_identical_windows(ref_over_shim, golden):
    """Predicted == observed: the reference raises TargetSourceCDFError from wasser(checkCommonCDF=True)
    (libs/OTlib.py:663-666 via MargWasserstein :1111-1113); so does the shim."""
    ru, OT = ref_over_shim
    g = golden("ricker_forward")
    grid = tuple(g["grid"][:4]) + (int(g["grid"][4]), int(g["grid"][5]))
    lam = float(g["lam"])
    to, wo = ru.rickerwavelet(0.0, 1.6, 1.0, trange=[-2.0, 2.0])
    wfo, tgt = ru.BuildOTobjfromWaveform(to, wo, grid, lambdav=lam)
    ru.ricker_util_opt.init()
    with pytest.raises(OT.TargetSourceCDFError):
        ru.optfunc(np.array([0.0, 1.6, 1.0]), [tgt, "W2", [-2.0, 2.0], grid, lam, False, 0.5, 45.0])
