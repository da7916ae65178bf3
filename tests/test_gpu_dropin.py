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


@pytest.mark.parametrize("transform", [False, True])
def test_unmodified_fd_checkers_over_shim(ref_over_shim, transform):
    """The reference's own finite-difference checkers (libs/ricker_util.py:554-606), unmodified, running over the shim:
    the walk-through of Ricker_waveform_derivatives.ipynb (cells 7-15, 31-50).  Analytic derivatives -
    MargWasserstein(derivatives, returnmargW) -> PDFderivMarg -> (arctan chain) -> dudm.dot - against central differences
    of the misfit itself: a check that needs no oracle.  The notebook's own table agrees to ~7 digits for dW/du and
    for the origin-time parameter, and to 3-4 digits for the amplitude marginal's dW/dm (its FD step is coarse)."""
    ru, OT = ref_over_shim
    trange = [-2.0, 2.0]
    mstart = np.array([5.0, 3.0, 0.5])                                             # cell 7
    tpred, wpred, dudm = ru.rickerwavelet(mstart[0], mstart[1], mstart[2], trange=trange, deriv=True)   # cell 44
    tobs, wobs = ru.rickerwavelet(0.0, 1.6, 1.0, trange=trange)                    # noiseless observed wavelet
    lam, theta = 0.03, 45.0                                                        # cell 12
    grid = (trange[0], trange[1], -0.8, 1.8, 80, 512) if transform else (trange[0], trange[1], -2.0, 3.5, 80, 512)
    wfobs, tgt = ru.BuildOTobjfromWaveform(tobs, wobs, grid, lambdav=lam, transform=transform, theta=theta)   # cell 15
    wfpred, src = ru.BuildOTobjfromWaveform(tpred, wpred, grid, lambdav=lam, deriv=True, transform=transform, theta=theta)
    w, dwdpbar, dwdt0 = OT.MargWasserstein(src, tgt, derivatives=True, distfunc="W2", returnmargW=True)   # cell 31
    wfpred.PDFderivMarg(dwdpbar)                                                   # cell 38
    gt, gu = wfpred.pdfdMarg[0].copy(), wfpred.pdfdMarg[1].copy()
    if transform:
        un, dundu = ru.arctan_trans(wpred, grid[2], grid[3], deriv=True)
        gt, gu = gt * dundu, gu * dundu
    # cell 41: d(Wt, Wu)/du at waveform points, 0.001 % central differences
    scale_t, scale_u = np.abs(gt).max(), np.abs(gu).max()
    for k in (23, 29, 44, 103, 177):
        fdt, fdu = ru.check_dwduFD(k, tpred, wpred, 0.001, grid, lam, tgt, transform=transform, theta=theta)
        assert fdt == pytest.approx(gt[k], rel=2e-5, abs=2e-7 * scale_t)
        assert fdu == pytest.approx(gu[k], rel=2e-5, abs=2e-7 * scale_u)
    # cells 48-50: d(Wt, Wu)/d(t0, A, f)
    dwtdm, dwudm = dudm.dot(gt), dudm.dot(gu)
    dwtdm[0] = dwdt0[0] / (wfpred.tant * (wfpred.tlim[1] - wfpred.tlim[0]))
    dwudm[0] = dwdt0[1] / (wfpred.tant * (wfpred.tlim[1] - wfpred.tlim[0]))
    fdt, fdu = ru.check_dwdmFD(0, tpred, wpred, 0.00001, mstart, grid, lam, tgt, trange, transform=transform, theta=theta)
    assert fdt == pytest.approx(dwtdm[0], rel=1e-6)
    assert abs(fdu) <= 1e-9 and dwudm[0] == 0.0
    for k in (1, 2):
        fdt, fdu = ru.check_dwdmFD(k, tpred, wpred, 0.00001, mstart, grid, lam, tgt, trange, transform=transform, theta=theta)
        assert fdu == pytest.approx(dwudm[k], rel=5e-3)


@pytest.mark.parametrize("cmt", [False, True])
def test_unmodified_loc_cmt_util_optfunc_OT_over_shim(ref_over_shim, golden, cmt):
    """The reference's CMT misfit function, UNMODIFIED (libs/loc_cmt_util.py:186-306 optfunc_OT with its own
    BuildOTobjfromWaveform / CalcWasserWaveform / arctan_trans / buildFingerprintwindows, :430-587), running over the
    shim: 4 stations x 3 components x 61 samples, 79 x 61 grids, arctan transform, source location only (3 parameters)
    and location + moment tensor (9), Wopt = Wavg / Wt / Wu and both marginals.  Expected values:
    tests/golden/cmt_optfunc.npz, produced by the same calls on the unmodified reference end to end
    (make_golden.py cmt).  pyprop8 is absent from the image: oracle/pyprop8_stub.py generates the seismograms and
    Jacobians on both sides (the fixture holds them too, as a check that both sides saw the same input)."""
    ru, OT = ref_over_shim
    from oracle import build_ref, cmt_scenario
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libs", "loc_cmt_util.py")):
        pytest.skip("oracle/_ref predates the CMT modules (python oracle/build_ref.py in the build container)")
    cmtu = build_ref.import_cmt()
    assert cmtu.OT is OT and os.path.realpath(cmtu.__file__).startswith(os.path.realpath(os.path.join(ROOT, "oracle", "_ref")))
    g = golden("cmt_optfunc")
    tag = "cmt" if cmt else "loc"
    optdata, t = cmt_scenario.build_optdata(cmtu, cmt=cmt)
    np.testing.assert_allclose(optdata["prop8data"]["obs_seis"], g[tag + "_obs"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(np.array(optdata["OTdata"]["obs_grids"], dtype=np.float64), g[tag + "_grids"], rtol=1e-12)
    models = cmt_scenario.trial_models(cmtu, cmt=cmt)
    np.testing.assert_allclose(np.array(models), g[tag + "_models"], rtol=1e-14)
    for i, m in enumerate(models):
        mis, dmis, tt, seis = cmtu.optfunc_OT(m, optdata, returnseis=True)
        np.testing.assert_allclose(seis, g[tag + "_seis"][i], rtol=1e-12, atol=1e-15)
        assert mis == pytest.approx(float(g[tag + "_mis"][i]), rel=1e-9)
        scale = np.abs(g[tag + "_dmis"][i]).max()
        np.testing.assert_allclose(dmis, g[tag + "_dmis"][i], rtol=1e-6, atol=1e-8 * scale)
    mis, dmis = cmtu.optfunc_OT(models[0], optdata, return2W=True)
    np.testing.assert_allclose(mis, g[tag + "_mis2W"], rtol=1e-9)
    np.testing.assert_allclose(np.array(dmis), g[tag + "_dmis2W"], rtol=1e-6, atol=1e-8 * np.abs(g[tag + "_dmis2W"]).max())
    for w in ("Wt", "Wu"):
        optdata["OTdata"]["Wopt"] = w
        mis, dmis = cmtu.optfunc_OT(models[1], optdata)
        assert mis == pytest.approx(float(g[tag + "_mis" + w]), rel=1e-9)
        np.testing.assert_allclose(dmis, g[tag + "_dmis" + w], rtol=1e-6, atol=1e-8 * np.abs(g[tag + "_dmis" + w]).max())
    assert len(cmtu.loc_cmt_util_opt.opt_history_data) == len(models) + 3          # the reference's history side effect


def test_out_of_scope_names_are_served_by_the_reference(ref_over_shim, golden):
    """With the shim installed over the reference package, names outside the accelerated path resolve to the
    reference's own functions (adapters.reference_attr) and work on shim objects through the attribute protocol:
    the reference's host-side evaluator fp.wavedistv (libs/FingerprintLib.py:456-474) applied to a shim waveformFP
    reproduces the GPU distance field, nearest segments and ray parameters bit for bit."""
    ru, OT = ref_over_shim
    fp = ru.fp
    assert fp.wavedistv.__module__ == "libs._reference_FingerprintLib"
    assert OT.wasserNumInt.__module__ == "libs._reference_OTlib"
    assert not hasattr(OT, "no_such_name")
    g = golden("ricker_forward")
    grid = tuple(g["grid"][:4]) + (int(g["grid"][4]), int(g["grid"][5]))
    wf = fp.waveformFP(g["to"], g["wo"], grid)
    wf.calcpdf(lambdav=float(g["lam"]))
    Xn, Yn = np.meshgrid(np.linspace(wf.tlimnfp[0], wf.tlimnfp[1], wf.ntg), np.linspace(wf.ulimnfp[0], wf.ulimnfp[1], wf.nug))
    points = np.vstack((Xn.flatten(), Yn.flatten())).T
    ks = np.random.default_rng(0).choice(points.shape[0], 4000, replace=False)
    d, irays, xrays, lrays = fp.wavedistv(points[ks], wf)            # the reference's NumPy evaluator, shim geometry
    np.testing.assert_array_equal(irays, wf.irays[ks])
    np.testing.assert_array_equal(d, wf.dfield.reshape(-1)[ks])
    np.testing.assert_array_equal(lrays, wf.lrays[ks])
    np.testing.assert_array_equal(xrays, wf.xrays[ks])


def _run_notebook(name, ns=None):
    """Execute the code cells of an UNMODIFIED reference notebook (oracle/_ref/notebooks, copied by oracle/build_ref.py)
    in one namespace, as Jupyter would, except for IPython magics and cells that draw (plt.* / *plot*( calls: matplotlib is
    not in the image).  Returns {cell index: captured stdout}."""
    import contextlib
    import io
    import json
    import re
    path = os.path.join(ROOT, "oracle", "_ref", "notebooks", name)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/notebooks not built (python oracle/build_ref.py in the build container)")
    ns = {} if ns is None else ns
    out = {}
    for i, c in enumerate(json.load(open(path))["cells"]):
        if c["cell_type"] != "code":
            continue
        src = "".join(l for l in c["source"] if not l.lstrip().startswith(("%", "!")))
        if re.search(r"\bplt\.\w+\(|\w*plot\w*\(", src):
            continue
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            exec(compile(src, "%s[cell %d]" % (name, i), "exec"), ns)
        out[i] = buf.getvalue()
    return out


def _table_rows(text, ncols):
    """Rows of a printed table: lines that consist of exactly `ncols` numbers."""
    rows = []
    for line in text.splitlines():
        tok = line.split()
        try:
            vals = [float(x) for x in tok]
        except ValueError:
            continue
        if len(vals) == ncols:
            rows.append(vals)
    return np.array(rows)


def test_point_mass_notebook_runs_unchanged(ref_over_shim):
    """Point_mass_demo_Fig_5.ipynb, code cells as they are, over the shim: the printed W_1 / W_2 are the notebook's."""
    out = _run_notebook("Point_mass_demo_Fig_5.ipynb")
    assert out[11].split() == ["W_1", "=", "4.11"]              # the notebook's stored outputs
    assert out[13].split() == ["W_2", "=", "18.09"]


def test_ricker_derivatives_notebook_runs_unchanged(ref_over_shim):
    """Ricker_waveform_derivatives.ipynb, code cells as they are (fingerprints, MargWasserstein, PDFderivMarg, the three
    finite-difference comparisons through fp.check_FDderiv, OT._checkderivMarg, ru.check_dwduFD / check_dwdmFD), over the
    shim.  The notebook's observed waveform carries Gaussian-process noise whose draw depends on the installed
    scikit-learn, so its stored digits are not reproducible anywhere; what the notebook demonstrates - analytic and
    finite-difference derivatives side by side - is checked on the tables it prints."""
    np.random.seed(12345)                                       # the notebook picks its comparison points with np.random
    out = _run_notebook("Ricker_waveform_derivatives.ipynb")
    t31 = _table_rows(out[31], 6)                               # grid point, segment, dd/du0, dd/du1 (FD), same (analytic)
    assert len(t31) == 30
    ok = np.abs(t31[:, 2:4] - t31[:, 4:6]).max(axis=1) <= 2e-2 * np.abs(t31[:, 4:6]).max(axis=1) + 2e-5
    assert ok.mean() >= 0.8                                     # the FD is off where the step changes the nearest segment (:517)
    t36 = _table_rows(out[36], 5)                               # grid point, dWt/dp, dWu/dp (FD), same (analytic)
    assert len(t36) >= 20
    np.testing.assert_allclose(t36[:, 1:3], t36[:, 3:5], rtol=1e-4, atol=2e-9)
    t41 = _table_rows(out[41], 5)                               # waveform point, dWt/du, dWu/du (FD), same (analytic)
    assert len(t41) == 10
    np.testing.assert_allclose(t41[:, 1:3], t41[:, 3:5], rtol=1e-3, atol=2e-9)
    rows50 = [l.split() for l in out[50].splitlines() if "parameter" in l and len(l.split()) >= 6][-3:]
    fd_t0, an_t0 = float(rows50[0][-4]), float(rows50[0][-2])   # time offset: FD and analytic dWt/dm
    assert fd_t0 == pytest.approx(an_t0, rel=1e-6) and abs(an_t0) > 0.1
    assert float(rows50[1][-3]) == pytest.approx(float(rows50[1][-1]), rel=5e-3)   # amplitude parameter, dWu/dm
    assert float(rows50[2][-3]) == pytest.approx(float(rows50[2][-1]), rel=5e-3)   # frequency parameter, dWu/dm


def test_ricker_inversion_notebook_runs_unchanged(ref_over_shim, golden):
    """Ricker_Figs_3_8.ipynb, code cells as they are, over the shim: the notebook's L-BFGS-B inversion
    `minimize(ru.optfunc, mstart, data, jac=True, ...)` (cell 32) of the double Ricker wavelet's (time offset, amplitude,
    frequency factor) from a noisy observation - the inversion loop this library exists to drop into.  The Gaussian-process
    noise draw depends on the installed scikit-learn, so the notebook's stored digits are not reproducible; the run must
    converge from the notebook's start (5.0, 3.0, 0.5) to its true model (0, 1.6, 1) within the noise."""
    ns = {}
    out = _run_notebook("Ricker_Figs_3_8.ipynb", ns)
    opt1, mtrue, mstart = ns["opt1"], ns["mtrue"], ns["mstart"]
    assert opt1.nfev >= 5 and len(ns["ricker_util_opt"].Wdata) == opt1.nfev      # every evaluation went through ru.optfunc
    w0, _ = ns["ru"].optfunc(mstart, ns["data"])
    assert opt1.fun < 1e-2 * w0                                                  # the misfit dropped by orders of magnitude
    np.testing.assert_allclose(opt1.x, mtrue, atol=0.2)
    assert len(ns["was"]) == len(ns["ls"]) >= 3                                  # cells 36-38: the per-iteration history
    # the same notebook run on the unmodified reference in the build container (make_golden.py inversion): if this box
    # drew the same noise, the optimiser must walk the same path to the same point
    g = golden("inversion_notebook")
    same = ns["wobs"].shape == g["wobs"].shape and np.allclose(ns["wobs"], g["wobs"], rtol=1e-12, atol=1e-14)
    print("inversion notebook: same noise draw as in the build container:", same, "x =", opt1.x, "nfev =", opt1.nfev)
    if same:
        assert opt1.nfev == int(g["nfev"]) and opt1.nit == int(g["nit"])
        np.testing.assert_allclose(opt1.x, g["x"], rtol=1e-6, atol=1e-8)
        assert opt1.fun == pytest.approx(float(g["fun"]), rel=1e-6)
        np.testing.assert_allclose(np.array(ns["ricker_util_opt"].Wits), g["its"], rtol=1e-6, atol=1e-8)
