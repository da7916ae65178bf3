"""Pin oracle/wfot_oracle.py against (a) the known answers printed in the
reference notebooks and (b) fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import wfot_oracle as O

WINDOW_CASES = ["small_q1", "small_q2", "small_theta", "small_fpgrid", "cmt_window"]


def _pair(g):
    q = None if int(g["q"]) < 0 else int(g["q"])
    fpgrid = tuple(g["fpgrid"]) if g["fpgrid"].size else None
    grid = tuple(g["grid"][:4]) + (int(g["grid"][4]), int(g["grid"][5]))
    lam, theta, distfunc = float(g["lam"]), float(g["theta"]), str(g["distfunc"])
    win = O.make_window(g["tp"], g["wp"], grid, fpgrid=fpgrid, theta=theta)
    O.calcpdf(win, q=q, lambdav=lam, deriv=True)
    wino = O.make_window(g["to"], g["wo"], grid, fpgrid=fpgrid, theta=theta)
    O.calcpdf(wino, q=q, lambdav=lam, deriv=False)
    src, tgt = O.otpdf(win.pdf, win.pos), O.otpdf(wino.pdf, wino.pos)
    return win, src, tgt, distfunc


# ---- known answers printed in the reference notebooks --------------------

def test_kat_point_mass_demo():
    # Point_mass_demo_Fig_5.ipynb cells 3, 11, 13: W1 = 4.11, W2^2 = 18.09
    fx, gx = np.linspace(3, 14, 6), np.linspace(7, 18, 6)
    f = np.array([.2, .01, .18, .21, .2, .2])
    g = np.array([.18, .07, .2, .05, .27, .23])
    out = O.wasser(O.otpdf(f, fx), O.otpdf(g, gx), "W12", derivatives=True)
    assert abs(out[0] - 4.11) < 1e-12 and abs(out[3] - 18.09) < 1e-12
    np.testing.assert_allclose(out[1], [6.16, 3.96, 1.76, -0.44, -2.64, -4.84], atol=1e-12)
    np.testing.assert_allclose(out[4], [49.28, 26.84, 14.08, 1.32, -21.12, -43.56], atol=1e-11)
    assert abs(out[2] + 1.0) < 1e-14 and abs(out[5] + 8.22) < 1e-12


def test_kat_ricker_derivatives_notebook():
    # Ricker_waveform_derivatives.ipynb cells 7/12/14/23/24/31 (noise-free predicted waveform)
    tp, wp = O.rickerwavelet(5.0, 3.0, 0.5, trange=(-2, 2))
    win = O.make_window(tp, wp, (-2, 2, -2.0, 3.5, 80, 512))
    O.calcpdf(win, lambdav=0.03, deriv=True)
    assert win.dfield.shape == (80, 512)
    np.testing.assert_allclose(win.dfield[0, :3], [0.17128621, 0.16994805, 0.16862197], atol=5e-9)
    np.testing.assert_allclose(win.dfield[-1, -3:], [0.25586098, 0.25766182, 0.25946493], atol=5e-9)
    np.testing.assert_allclose(win.pdf[0, :3], [0.0033142, 0.00346537, 0.00362199], atol=5e-9)
    assert list(win.irays[:5]) == [29, 29, 29, 29, 29]
    np.testing.assert_allclose(win.dddy[:3], [[0, 0.13214619], [0, 0.13318671], [0, 0.13423411]], atol=5e-9)
    # SURVEY appendix B extra known answers
    assert abs(win.pdf.sum() - 9367.6) < 0.05
    np.testing.assert_allclose(win.pdf.sum(axis=0)[:3], [6.1175, 6.2894, 6.4536], atol=5e-5)
    np.testing.assert_allclose(win.pdf.sum(axis=1)[:3], [5.7070, 8.4160, 12.3549], atol=5e-5)


# ---- fixtures produced by the unmodified reference -----------------------

def test_pointmass_fixture(golden):
    g = golden("pointmass")
    s, t = O.otpdf(g["f"], g["fx"]), O.otpdf(g["g"], g["gx"])
    np.testing.assert_array_equal(s.cdf, g["cdf_f"])
    np.testing.assert_array_equal(t.cdf, g["cdf_g"])
    out = O.wasser(s, t, "W12", derivatives=True)
    for got, key in zip(out, ["W1", "dW1", "dW1pos", "W2", "dW2", "dW2pos"]):
        np.testing.assert_allclose(got, g[key], rtol=0, atol=1e-13)


def test_ot1d_random_fixture(golden):
    g = golden("ot1d_random")
    s, t = O.otpdf(g["f"], g["xf"]), O.otpdf(g["g"], g["xg"])
    out = O.wasser(s, t, "W12", derivatives=True)
    for got, key in zip(out, ["W1", "dW1", "dW1pos", "W2", "dW2", "dW2pos"]):
        np.testing.assert_allclose(got, g[key], rtol=1e-13, atol=1e-15)
    lin = O.wasser_linear(s, t, "W12")           # O(n) derivative identity, SURVEY A.6
    for got, key in zip(lin, ["W1", "dW1", "dW1pos", "W2", "dW2", "dW2pos"]):
        np.testing.assert_allclose(got, g[key], rtol=1e-10, atol=1e-13)
    s2, t2 = O.otpdf(g["f2"], g["x2f"]), O.otpdf(g["g2"], g["x2g"])   # n != m, no derivatives
    out2 = O.wasser(s2, t2, "W12")
    np.testing.assert_allclose(out2, [g["W1_2"], g["W2_2"]], rtol=1e-13)


@pytest.mark.parametrize("case", WINDOW_CASES)
def test_window_fixture_all_fields(golden, case):
    g = golden(case)
    win, src, tgt, distfunc = _pair(g)
    np.testing.assert_array_equal(win.pn, g["pn"])                # bit-exact normalisation
    np.testing.assert_array_equal(win.lsq_n, g["lsq_n"])
    np.testing.assert_array_equal(win.irays, g["irays"])          # bit-exact nearest segment
    np.testing.assert_array_equal(win.lrays, g["lrays"])
    np.testing.assert_array_equal(win.dfield, g["dfield"])
    np.testing.assert_array_equal(win.xrays, g["xrays"])
    np.testing.assert_allclose(win.pdf, g["pdf"], rtol=1e-15)
    np.testing.assert_allclose(win.dddy, g["dddy"], rtol=1e-12, atol=1e-15)
    assert src.amp == pytest.approx(float(g["amp"]), rel=1e-15)
    np.testing.assert_allclose(src.pdf.sum(axis=0) / src.pdf.sum(), g["marg_t"], rtol=1e-14)
    W, dW, dwg = O.marg_wasserstein(src, tgt, distfunc=distfunc, derivatives=True, returnmargW=True)
    np.testing.assert_array_equal(src.marg[0].cdf, g["cdf_t"])
    np.testing.assert_array_equal(src.marg[1].cdf, g["cdf_u"])
    np.testing.assert_array_equal(tgt.marg[0].cdf, g["tgt_cdf_t"])
    np.testing.assert_allclose(W, g["W"], rtol=1e-13)
    np.testing.assert_allclose(dwg, g["dwg"], rtol=1e-13)
    np.testing.assert_allclose(dW[0], g["dWt"], rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(dW[1], g["dWu"], rtol=1e-11, atol=1e-15)
    pm = O.pdfderiv_marg(win, dW)
    np.testing.assert_allclose(pm[0], g["pdfdMarg0"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(pm[1], g["pdfdMarg1"], rtol=1e-10, atol=1e-14)
    Wavg, dWavg, dwgavg = O.marg_wasserstein(src, tgt, distfunc=distfunc, derivatives=True)
    assert Wavg == pytest.approx(float(g["Wavg"]), rel=1e-13)
    assert dwgavg == pytest.approx(float(g["dwgavg"]), rel=1e-13)
    np.testing.assert_allclose(O.pdfderiv(win, dWavg), g["pdfd"], rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("case", ["ricker_cfg1", "ricker_cfg1_w1"])
def test_ricker_cfg1_fixture(golden, case):
    g = golden(case)
    win, src, tgt, distfunc = _pair(g)
    k = np.arange(0, win.irays.size, int(g["sub"]))
    np.testing.assert_array_equal(win.irays, g["irays_all"].astype(np.int64))
    np.testing.assert_array_equal(win.dfield.reshape(-1)[k], g["dfield_sub"])
    np.testing.assert_array_equal(win.lrays[k], g["lrays_sub"])
    np.testing.assert_allclose(win.pdf.reshape(-1)[k], g["pdf_sub"], rtol=1e-15)
    np.testing.assert_allclose(win.dddy[k], g["dddy_sub"], rtol=1e-11, atol=1e-15)
    assert win.dfield.sum() == pytest.approx(float(g["sum_dfield"]), rel=1e-14)
    assert win.pdf.sum() == pytest.approx(float(g["sum_pdf"]), rel=1e-14)
    W, dW, dwg = O.marg_wasserstein(src, tgt, distfunc=distfunc, derivatives=True, returnmargW=True)
    np.testing.assert_allclose(W, g["W"], rtol=1e-13)
    np.testing.assert_allclose(dwg, g["dwg"], rtol=1e-13)
    np.testing.assert_allclose(dW[0][0, :], g["dWt_row0"], rtol=1e-10, atol=1e-15)
    np.testing.assert_allclose(dW[1][:, 0], g["dWu_col0"], rtol=1e-10, atol=1e-15)
    pm = O.pdfderiv_marg(win, dW)
    np.testing.assert_allclose(pm[0], g["pdfdMarg0"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(pm[1], g["pdfdMarg1"], rtol=1e-9, atol=1e-13)


def test_error_behaviour():
    x = np.linspace(0, 1, 5)
    with pytest.raises(O.PDFSignError):
        O.otpdf(np.array([0.1, -0.2, 0.3, 0.2, 0.2]), x)
    with pytest.raises(O.PDFShapeError):
        O.otpdf(np.ones(4), x)
    p = O.otpdf(np.ones(5), x)
    with pytest.raises(O.TargetSourceCDFError):      # identical CDFs: SURVEY appendix B
        O.wasser(p, O.otpdf(np.ones(5), x), "W2", derivatives=True)
    with pytest.raises(O.TargetSource2DShapeError):
        O.marg_wasserstein(p, p)


def test_ricker_forward_and_optfunc_fixture(golden):
    """Forward model + parameter derivatives bit-for-bit, optfunc chain to 1e-12 (libs/ricker_util.py:38-89,373-404)."""
    g = golden("ricker_forward")
    for i, (tp, a, f) in enumerate(g["params"]):
        t, w, dw = O.rickerwavelet(tp, a, f, deriv=True)
        np.testing.assert_array_equal(t, g["t"][i])
        np.testing.assert_array_equal(w, g["w"][i])
        np.testing.assert_array_equal(dw, g["dw"][i])
    grid = tuple(g["grid"][:4]) + (int(g["grid"][4]), int(g["grid"][5]))
    _, tg = O.build_ot_from_waveform(g["to"], g["wo"], grid, lambdav=float(g["lam"]))
    f_, g_ = O.ricker_optfunc(g["X"][0], tg, "W2", (-2.0, 2.0), grid, float(g["lam"]), float(g["alpha"]))
    assert f_ == pytest.approx(float(g["F"][0]), rel=1e-12)
    np.testing.assert_allclose(g_, g["G"][0], rtol=1e-10, atol=1e-14)


def test_sliced_wasserstein_and_plan_fixture(golden):
    """Sliced Wasserstein (libs/OTlib.py:119-144,1156-1318) and the transport plan of wasser (:718-740)."""
    g = golden("sliced_plan")
    for d in ("W1", "W2"):
        r = O.sliced_wasserstein(O.otpdf(g["f"], g["pos"]), O.otpdf(g["g"], g["pos"]), 6, d, derivatives=True)
        assert r[0] == pytest.approx(float(g["sw_" + d]), rel=1e-13)
        np.testing.assert_allclose(r[1], g["dsw_" + d], rtol=1e-11, atol=1e-15)
    w = O.wasser(O.otpdf(g["plan_f"], g["plan_fx"]), O.otpdf(g["plan_g"], g["plan_gx"]), "W2", derivatives=True,
                 returnplan=True)
    np.testing.assert_allclose(w[3], g["plan_H"], atol=1e-15)
    np.testing.assert_allclose(w[4], g["plan_dH"], atol=1e-14)


@pytest.mark.parametrize("tag,m", [("loc", 0), ("cmt", 2)])
def test_cmt_optfunc_fixture(golden, tag, m):
    """The oracle's CMT adapter composition (arctan window per station/component -> fingerprint -> marginal W2 -> chain
    to the seismogram and on to the model parameters) against what the UNMODIFIED libs/loc_cmt_util.optfunc_OT
    (:186-306) returned on the unmodified reference: tests/golden/cmt_optfunc.npz (make_golden.py cmt)."""
    g = golden("cmt_optfunc")
    seis, J, obs, grids = g[tag + "_seis"][m], g[tag + "_J"][m], g[tag + "_obs"], g[tag + "_grids"]
    nr, nc, nt = seis.shape
    t = np.arange(float(nt))
    tot, dr = 0.0, np.zeros((nr, nc, nt))
    for i in range(nr):
        for j in range(nc):
            grid = tuple(grids[i, j][:4]) + (int(grids[i, j][4]), int(grids[i, j][5]))
            assert list(grid) == list(O.build_fingerprint_window(t, obs[i, j]))                    # :430-446
            _, tgt = O.build_ot_from_waveform(t, obs[i, j], grid, lambdav=0.04, transform=True)
            W, d, dg, _, _ = O.misfit_grad_window(t, seis[i, j], grid, tgt, lambdav=0.04, transform=True, adapter="cmt")
            tot += 0.5 * (W[0] + W[1])                                                             # Wopt = 'Wavg'
            dr[i, j] = 0.5 * (d[0] + d[1])
    assert tot == pytest.approx(float(g[tag + "_mis"][m]), rel=1e-12)
    np.testing.assert_allclose(dr, g[tag + "_dr"][m], rtol=1e-9, atol=1e-12 * np.abs(dr).max())
    np.testing.assert_allclose(J.dot(dr.reshape(-1)), g[tag + "_dmis"][m], rtol=1e-9, atol=1e-12 * np.abs(g[tag + "_dmis"][m]).max())


def test_inversion_notebook_fixture(golden):
    """The oracle's optfunc composition inside the notebook's own optimiser call (Ricker_Figs_3_8.ipynb cells 11-32:
    L-BFGS-B from (5, 3, 0.5), jac=True, tol 1e-8) walks the path the unmodified reference walked on the same noisy
    observation: tests/golden/inversion_notebook.npz (make_golden.py inversion)."""
    from scipy.optimize import minimize
    g = golden("inversion_notebook")
    trange, lam, alpha = [-2.0, 2.0], 0.03, 0.5
    grid = (trange[0], trange[1], -2.0, 3.5, 80, 512)                              # cells 14, 17
    tobs, _ = O.rickerwavelet(0.0, 1.6, 1.0, trange=trange)                         # the time axis; the amplitudes (with the
    _, tgt = O.build_ot_from_waveform(tobs, g["wobs"], grid, lambdav=lam)           # reference's noise) come from the fixture
    its = [[5.0, 3.0, 0.5]]
    opt = minimize(lambda x: O.ricker_optfunc(x, tgt, "W2", trange, grid, lam, alpha=alpha), np.array(its[0]),
                   jac=True, tol=1e-8, method="L-BFGS-B", options={"maxiter": 500}, callback=lambda x: its.append(x.tolist()))
    assert opt.nfev == int(g["nfev"]) and opt.nit == int(g["nit"])
    np.testing.assert_allclose(opt.x, g["x"], rtol=1e-6, atol=1e-8)
    assert opt.fun == pytest.approx(float(g["fun"]), rel=1e-6)
    np.testing.assert_allclose(np.array(its), g["its"], rtol=1e-6, atol=1e-8)


def test_fd_checkers_fixture(golden):
    """The oracle's analytic derivatives against the finite differences the reference's own checkers produced
    (fp.check_FDderiv, OT._checkderivMarg: tests/golden/fd_checkers.npz, make_golden.py fd) and against the
    reference's analytic dddy at the same grid points."""
    g = golden("fd_checkers")
    grid = tuple(g["grid"][:4]) + (int(g["grid"][4]), int(g["grid"][5]))
    lam = float(g["lam"])
    _, tgt = O.build_ot_from_waveform(g["to"], g["wo"], grid, lambdav=lam)
    win, src = O.build_ot_from_waveform(g["tp"], g["wp"], grid, lambdav=lam, deriv=True)
    ks = g["ks"].astype(int)
    np.testing.assert_array_equal(win.irays[ks], g["fd"][:, 0].astype(int))
    np.testing.assert_allclose(win.dddy[ks], g["dddy"], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(win.dddy[ks], g["fd"][:, 1:], rtol=2e-2, atol=1e-5)        # the reference's FD step: ~1 %
    out = O.marg_wasserstein(src, tgt, distfunc="W2", derivatives=True, returnmargW=True)
    dWt, dWu = out[1]
    np.testing.assert_allclose(np.stack([dWt.reshape(-1)[ks], dWu.reshape(-1)[ks]], axis=1), g["marg"], rtol=1e-4, atol=1e-9)
