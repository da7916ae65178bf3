"""CPU: host-side logic of the reference-facing shim (no GPU work)."""
import pickle

import numpy as np
import pytest

from oracle import wfot_oracle as O


def test_grid_struct_layout():
    from waveform_ot_b200 import _cabi, batch
    assert batch.GRID_DTYPE.itemsize == 80
    for name, (dt, off) in batch.GRID_DTYPE.fields.items():
        assert getattr(_cabi.wfot_grid, name).offset == off


def test_no_cpu_fallback():
    """The product path must fail loudly without a CUDA device."""
    import torch
    from waveform_ot_b200 import batch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        batch.fingerprint_batch(np.linspace(0, 1, 8), np.zeros(8), (0, 1, -1, 1, 4, 4), 4, 4, 0.04)
    import waveform_ot_b200.batch as b
    src = open(b.__file__).read() + open(b.__file__.replace("batch.py", "OTlib.py")).read()
    assert "oracle" not in src              # the shim never imports the oracle


@pytest.mark.parametrize("kw", [dict(), dict(theta=60.0), dict(tantheta=0.5),
                                dict(fpgrid=(0.2, 3.9, -2.0, 2.2))])
def test_waveformfp_constructor_matches_reference_semantics(kw):
    """waveformFP.__init__ is host-only: same attributes as libs/FingerprintLib.py:75-115."""
    from waveform_ot_b200 import FingerprintLib as fp
    rng = np.random.default_rng(1)
    t = np.sort(rng.random(30)) * 3 + 0.5
    w = rng.standard_normal(30)
    grid = (0.0, 4.0, -2.5, 2.5, 20, 16)
    wf = fp.waveformFP(t, w, grid, **kw)
    win = O.make_window(t, w, grid, **kw)
    assert (wf.ntg, wf.nug, wf.nt) == (16, 20, 30)
    assert wf.tant == pytest.approx(win.tant) and wf.theta == pytest.approx(win.theta)
    assert wf.tlimn == win.tlimn and wf.tlimnfp == win.tlimnfp and wf.ulimnfp == win.ulimnfp
    np.testing.assert_array_equal(wf.pn, win.pn)
    np.testing.assert_array_equal(wf.delta_n, win.delta_n)
    np.testing.assert_array_equal(wf.lsq_n, win.lsq_n)
    assert wf.x0.shape == (1, 29, 2) and not wf.dcalc
    with pytest.raises(fp.WaveformPFderivError):
        wf.wdistderiv()
    with pytest.raises(NotImplementedError):
        wf.calcpdf(method="FMM")
    with pytest.raises(fp.FingerprintMethodError):
        wf.calcpdf(method="bogus")
    wf2 = pickle.loads(pickle.dumps(wf))
    np.testing.assert_array_equal(wf2.pn, wf.pn)


def test_otlib_argument_errors():
    from waveform_ot_b200 import OTlib as OT
    with pytest.raises(OT.UnknownOTDistanceTypeError):
        OT._checkdistfunc(3.0)
    with pytest.raises(NotImplementedError):
        OT._checkdistfunc(np.zeros((2, 2)))
    assert OT._checkdistfunc("W12") == (True, True) and OT._checkdistfunc("W1") == (True, False)
    for name in ("PDFSignError", "PDFShapeError", "TargetSourceCDFError", "TargetSource2DShapeError",
                 "MarginalWassersteinError", "UnknownOTDistanceTypeError", "DistfuncShapeError"):
        assert issubclass(getattr(OT, name), Exception)


def test_adapters_arctan_and_install():
    from waveform_ot_b200 import adapters
    u = np.linspace(-3, 3, 11)
    a, da = adapters.arctan_trans(u, -1.0, 2.0, deriv=True)
    b, db = O.arctan_trans(u, -1.0, 2.0, deriv=True)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(da, db)
    fpm, otm = adapters.install("libs_test_prefix")
    import sys
    assert sys.modules["libs_test_prefix.FingerprintLib"] is fpm
    assert sys.modules["libs_test_prefix.OTlib"] is otm


def test_shard_bounds_cover_exactly():
    from waveform_ot_b200.dist import shard_bounds
    for n in (0, 1, 7, 30, 4096, 4_000_000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_build_fingerprint_windows_rule():
    """adapters.buildFingerprintwindows == libs/loc_cmt_util.py:429-446 (oracle restatement), no GPU needed."""
    import numpy as np
    from oracle import wfot_oracle as O
    from waveform_ot_b200 import adapters
    rng = np.random.default_rng(4)
    t = np.arange(61.0)
    wave = rng.standard_normal((3, 2, 61)) * 1e-3
    g = adapters.buildFingerprintwindows(t, wave)
    for i in range(3):
        for j in range(2):
            assert list(g[i][j]) == list(O.build_fingerprint_window(t, wave[i, j]))
    g2 = adapters.buildFingerprintwindows(t, wave, Nu=50, Nt=40, u0=-1.0, u1=2.0)
    assert g2[1][1][2:] == [-1.0, 2.0, 50, 40]


def test_shard_rows_periodic_rotation():
    """dist._shard_rows: shard-local window b must find the row of global window lo + b
    (the kernels index periodic rows with b % rows)."""
    import torch
    from waveform_ot_b200.dist import _shard_rows, shard_bounds
    n, rows = 900, 9
    x = torch.arange(rows * 2, dtype=torch.float64).reshape(rows, 2)
    per_window = torch.arange(n * 2, dtype=torch.float64).reshape(n, 2)
    for world in (1, 2, 7, 8):
        for rank in range(world):
            lo, hi = shard_bounds(n, rank, world)
            y = _shard_rows(x, n, lo, hi, "grids")
            for b in (0, 1, rows - 1, rows, hi - lo - 1):
                assert torch.equal(y[b % rows], x[(lo + b) % rows])
            z = _shard_rows(per_window, n, lo, hi, "grids")
            assert z.shape[0] == hi - lo and torch.equal(z[0], per_window[lo])
            assert _shard_rows(x[:1], n, lo, hi, "grids") is not None
    with pytest.raises(ValueError):
        _shard_rows(torch.zeros(7, 2), n, 5, 10, "grids")      # 7 rows do not tile 900 windows


def test_sharded_empty_shard_contributes_zeros(monkeypatch):
    """fewer windows than ranks: the empty rank must not call the kernel (it would raise INVALID_ARG while the
    other ranks wait in the allreduce) but add a zero vector."""
    import torch
    from waveform_ot_b200 import dist as wd
    monkeypatch.setenv("RANK", "3")
    monkeypatch.setenv("WORLD_SIZE", "4")
    w = torch.zeros((2, 16))
    out = wd.misfit_grad_sharded(torch.linspace(0, 1, 16), w, None, 4, 4, 0.04, None)
    assert out.shape == (3 + 2 * 16,) and float(out.abs().sum()) == 0.0


def test_model_chunk_bounds():
    """adapters._chunk_bounds: every model in exactly one chunk, no chunk larger than asked, a short first chunk."""
    from waveform_ot_b200 import adapters
    for M in (1, 2, 3, 5, 37, 255, 256, 257, 1024, 1025, 4096, 5000):
        for cm in (1, 2, 8, 64, 512, 1024, 10 ** 6):
            b = adapters._chunk_bounds(M, cm)
            assert b[0] == 0 and b[-1] == M and all(x < y for x, y in zip(b, b[1:]))
            assert max(y - x for x, y in zip(b, b[1:])) <= max(1, min(cm, M))
            if M > cm >= 4:
                assert b[1] == cm // 4
    assert adapters._chunk_bounds(4096, 1024) == [0, 256, 1280, 2304, 3328, 4096]


def test_out_of_scope_names_delegate_to_installed_reference():
    """Module __getattr__ of the shim: a clear AttributeError for names outside the path, and - once installed over a
    reference package - the reference's own function (no GPU needed: nothing is called)."""
    import importlib
    import sys
    from oracle import build_ref
    from waveform_ot_b200 import FingerprintLib, OTlib, adapters
    saved_prefix = adapters._installed_prefix
    adapters._installed_prefix = None
    try:
        with pytest.raises(AttributeError, match="outside the accelerated path"):
            OTlib.plotWasser
        with pytest.raises(AttributeError, match="has no attribute"):
            OTlib.no_such_name
        assert not hasattr(FingerprintLib, "plot_LS")
        if not build_ref.available():
            pytest.skip("oracle/_ref not built")
        saved = {k: v for k, v in sys.modules.items() if k == "libs" or k.startswith("libs.")}
        try:
            build_ref.import_reference()
            for k in [k for k in sys.modules if k == "libs" or k.startswith("libs.")]:
                del sys.modules[k]
            importlib.import_module("libs")
            adapters.install("libs")
            assert OTlib.wasserNumInt.__module__ == "libs._reference_OTlib"
            assert FingerprintLib.wavedistv.__module__ == "libs._reference_FingerprintLib"
            assert not hasattr(OTlib, "no_such_name")
        finally:
            for k in [k for k in sys.modules if k == "libs" or k.startswith("libs.")]:
                del sys.modules[k]
            sys.modules.update(saved)
    finally:
        adapters._installed_prefix = saved_prefix
