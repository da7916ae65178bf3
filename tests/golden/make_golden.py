"""Generate golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  The fixtures pin ``oracle/wfot_oracle.py`` and,
through it, the CUDA path.  Inputs are generated from fixed seeds and stored in
the fixture next to the outputs, so tests never need the reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refimport import import_reference  # noqa: E402

fp, OT, ru = import_reference()


def run_window(t, w, grid, lam, q=None, fpgrid=None, theta=45.0, tantheta=1.0):
    wf = fp.waveformFP(t, w, grid, fpgrid=fpgrid, theta=theta, tantheta=tantheta)
    wf.calcpdf(q=q, lambdav=lam, deriv=True)
    return wf


def full_pair(tag, tp, wp, to, wo, grid, lam, distfunc, q=None, fpgrid=None, theta=45.0,
              sub=None):
    """pred (tp,wp) vs obs (to,wo) on a common grid through the reference's
    own calls: waveformFP -> calcpdf -> OTpdf -> MargWasserstein -> PDFderivMarg."""
    wf = run_window(tp, wp, grid, lam, q=q, fpgrid=fpgrid, theta=theta)
    wo_ = run_window(to, wo, grid, lam, q=q, fpgrid=fpgrid, theta=theta)
    src = OT.OTpdf((wf.pdf, wf.pos))
    tgt = OT.OTpdf((wo_.pdf, wo_.pos))
    W, dW, dwg = OT.MargWasserstein(src, tgt, distfunc=distfunc, derivatives=True, returnmargW=True)
    wf.PDFderivMarg(dW)
    Wavg, dWavg, dwgavg = OT.MargWasserstein(src, tgt, distfunc=distfunc, derivatives=True)
    wf.PDFderiv(chainmatrix=dWavg)
    d = dict(
        tp=tp, wp=wp, to=to, wo=wo, grid=np.array(grid, dtype=np.float64), lam=lam,
        distfunc=distfunc, q=-1 if q is None else q,
        fpgrid=np.array(fpgrid if fpgrid is not None else [], dtype=np.float64), theta=theta,
        pn=wf.pn, lsq_n=wf.lsq_n, tlimn=np.array(wf.tlimn),
        amp=src.amp, marg_t=src.marg[0].pdf, marg_u=src.marg[1].pdf,
        cdf_t=src.marg[0].cdf, cdf_u=src.marg[1].cdf, x_t=src.marg[0].x, x_u=src.marg[1].x,
        tgt_amp=tgt.amp, tgt_cdf_t=tgt.marg[0].cdf, tgt_cdf_u=tgt.marg[1].cdf,
        tgt_x_t=tgt.marg[0].x, tgt_x_u=tgt.marg[1].x,
        W=np.array(W), dwg=np.array(dwg, dtype=np.float64),
        pdfdMarg0=wf.pdfdMarg[0], pdfdMarg1=wf.pdfdMarg[1],
        Wavg=Wavg, dwgavg=dwgavg, pdfd=wf.pdfd,
        sum_dfield=wf.dfield.sum(), sum_pdf=wf.pdf.sum(), sum_irays=int(wf.irays.sum()),
        sum_lrays=wf.lrays.sum(),
    )
    if sub is None:
        d.update(dfield=wf.dfield, irays=wf.irays.astype(np.int32), lrays=wf.lrays,
                 xrays=wf.xrays, pdf=wf.pdf, dddy=wf.dddy, dWt=dW[0], dWu=dW[1])
    else:  # strided subsample of the per-pixel fields (keeps the fixture small)
        k = np.arange(0, wf.irays.size, sub)
        d.update(sub=sub, dfield_sub=wf.dfield.reshape(-1)[k], irays_sub=wf.irays[k].astype(np.int32),
                 lrays_sub=wf.lrays[k], pdf_sub=wf.pdf.reshape(-1)[k], dddy_sub=wf.dddy[k],
                 irays_all=wf.irays.astype(np.int16),
                 dWt_row0=dW[0][0, :], dWu_col0=dW[1][:, 0])
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **d)
    print(tag, "W", W, "dwg", dwg)


def main():
    # ---- (1) point masses: Point_mass_demo_Fig_5.ipynb cells 3/11/13 ------
    fx = np.linspace(3, 14, 6)
    gx = np.linspace(7, 18, 6)
    f = np.array([.2, .01, .18, .21, .2, .2])
    g = np.array([.18, .07, .2, .05, .27, .23])
    s, t = OT.OTpdf((f, fx)), OT.OTpdf((g, gx))
    out = OT.wasser(s, t, 'W12', derivatives=True)
    np.savez(os.path.join(HERE, "pointmass.npz"), f=f, g=g, fx=fx, gx=gx,
             W1=out[0], dW1=out[1], dW1pos=out[2], W2=out[3], dW2=out[4], dW2pos=out[5],
             cdf_f=s.cdf, cdf_g=t.cdf)

    # ---- (2) 1-D random OT, n == m with derivatives, n != m without -------
    rng = np.random.default_rng(20221)
    n = 96
    f = rng.random(n) + 1e-3
    g = rng.random(n) + 1e-3
    x = np.linspace(0, 1, n)
    xg = np.sort(rng.random(n)) * 1.5 - 0.2
    s, t = OT.OTpdf((f, x)), OT.OTpdf((g, xg))
    out = OT.wasser(s, t, 'W12', derivatives=True)
    f2 = rng.random(50) + 1e-3
    g2 = rng.random(70) + 1e-3
    x2f, x2g = np.linspace(0, 1, 50), np.linspace(-0.1, 1.3, 70)
    s2, t2 = OT.OTpdf((f2, x2f)), OT.OTpdf((g2, x2g))
    out2 = OT.wasser(s2, t2, 'W12')
    np.savez(os.path.join(HERE, "ot1d_random.npz"), f=f, g=g, xf=x, xg=xg,
             W1=out[0], dW1=out[1], dW1pos=out[2], W2=out[3], dW2=out[4], dW2pos=out[5],
             f2=f2, g2=g2, x2f=x2f, x2g=x2g, W1_2=out2[0], W2_2=out2[1])

    # ---- (3) Ricker cfg1 shape, noise-free: Ricker_waveform_derivatives.ipynb cells 7/12/14
    tp, wp = ru.rickerwavelet(5.0, 3.0, 0.5, trange=[-2, 2])
    to, wo = ru.rickerwavelet(0.0, 1.6, 1.0, trange=[-2, 2])
    full_pair("ricker_cfg1", tp, wp, to, wo, (-2, 2, -2.0, 3.5, 80, 512), 0.03, "W2", sub=97)
    full_pair("ricker_cfg1_w1", tp, wp, to, wo, (-2, 2, -2.0, 3.5, 80, 512), 0.03, "W1", sub=97)

    # ---- (4) small random windows, all per-pixel fields kept --------------
    rng = np.random.default_rng(7)
    nt = 24
    t = np.sort(rng.random(nt)) * 3.0 + 0.5          # non-uniform sampling
    t[0], t[-1] = 0.5, 3.5
    wp = rng.standard_normal(nt).cumsum() * 0.3
    wo = wp + 0.25 * rng.standard_normal(nt)
    full_pair("small_q1", t, wp, t + 0.2, wo, (0.0, 4.0, -2.5, 2.5, 20, 16), 0.05, "W2")
    full_pair("small_q2", t, wp, t + 0.2, wo, (0.0, 4.0, -2.5, 2.5, 20, 16), 0.05, "W2", q=2)
    full_pair("small_theta", t, wp, t + 0.2, wo, (0.0, 4.0, -2.5, 2.5, 20, 16), 0.05, "W1", theta=60.0)
    full_pair("small_fpgrid", t, wp, t + 0.2, wo, (0.0, 4.0, -2.5, 2.5, 20, 16), 0.05, "W2",
              fpgrid=(0.2, 3.9, -2.0, 2.2))

    # ---- (5) cmt-shaped window: nt=61 -> 79x61, arctan transform, lambda=0.04
    rng = np.random.default_rng(11)
    t = np.arange(61.0)
    pulse = np.exp(-0.5 * ((t - 25.0) / 4.0) ** 2) * np.sin(0.5 * (t - 25.0))
    wp = 1e-3 * (pulse + 0.02 * rng.standard_normal(61))
    wo = 1e-3 * (np.roll(pulse, 3) * 1.2 + 0.03 * rng.standard_normal(61))
    du = wo.max() - wo.min()
    u0, u1 = wo.min() - 0.3 * du, wo.max() + 0.3 * du
    unp = ru.arctan_trans(wp, u0, u1)
    uno = ru.arctan_trans(wo, u0, u1)
    full_pair("cmt_window", t, unp, t, uno, (0.0, 60.0, 0.0, 1.0, 79, 61), 0.04, "W2")
    d = dict(np.load(os.path.join(HERE, "cmt_window.npz")))
    d.update(raw_wp=wp, raw_wo=wo, u0=u0, u1=u1)
    np.savez_compressed(os.path.join(HERE, "cmt_window.npz"), **d)

    # ---- (6) Ricker forward model + derivatives and the full optfunc chain
    #          (libs/ricker_util.py:38-89, 373-404; Ricker_Figs_3_8.ipynb cells 14/26/32 set-up, noise free)
    params = np.array([[0.0, 1.6, 1.0], [4.5, 1.6, 0.8], [5.0, 3.0, 0.5], [-1.3, 0.7, 1.4], [0.35, 2.2, 1.1]])
    T, Wv, DW = [], [], []
    for tp_, a_, f_ in params:
        t_, w_, dw_ = ru.rickerwavelet(tp_, a_, f_, trange=[-2., 2.], deriv=True)
        T.append(t_); Wv.append(w_); DW.append(dw_)
    grid = (-2, 2, -1.8, 4.2, 40, 128)
    lam, alpha = 0.03, 0.5
    to, wo = ru.rickerwavelet(0.0, 1.6, 1.0, trange=[-2., 2.])
    wfo, tgt = ru.BuildOTobjfromWaveform(to, wo, grid, lambdav=lam)
    X = np.array([[0.7, 1.3, 0.8], [-0.4, 2.0, 1.2], [1.5, 1.0, 0.9]])
    ru.ricker_util_opt.init()          # module-global history lists optfunc appends to (notebooks do this too)
    F, G = [], []
    for x in X:
        data = [tgt, 'W2', [-2., 2.], grid, lam, False, alpha, 45.0]
        w2, dv = ru.optfunc(x, data)
        F.append(w2); G.append(dv)
    np.savez_compressed(os.path.join(HERE, "ricker_forward.npz"), params=params, t=np.array(T), w=np.array(Wv),
                        dw=np.array(DW), grid=np.array(grid, dtype=np.float64), lam=lam, alpha=alpha, to=to, wo=wo,
                        X=X, F=np.array(F), G=np.array(G))
    print("ricker_forward", np.array(F), np.array(G)[0])

    # ---- (7) sliced Wasserstein (libs/OTlib.py:119-144,1156-1318) and the transport plan of wasser (:718-740)
    rng = np.random.default_rng(314)
    nx, ny = 7, 9
    X, Y = np.meshgrid(np.linspace(0, 1, ny), np.linspace(0, 1, nx))
    pos = np.stack([X, Y], axis=-1)
    f = rng.random((nx, ny)) + 0.05
    g = rng.random((nx, ny)) + 0.05
    out = {}
    for d in ("W1", "W2"):
        s, t = OT.OTpdf((f, pos)), OT.OTpdf((g, pos))
        r = OT.SlicedWasserstein(s, t, 6, distfunc=d, derivatives=True)
        out["sw_" + d], out["dsw_" + d] = r[0], r[1]
    s, t = OT.OTpdf((f, pos)), OT.OTpdf((g, pos))
    out["sw_noderiv"] = OT.SlicedWasserstein(s, t, 4, distfunc="W2")[0]
    fx, gx = np.linspace(3, 14, 6), np.linspace(7, 18, 6)
    f1 = np.array([.2, .01, .18, .21, .2, .2])
    g1 = np.array([.18, .07, .2, .05, .27, .23])
    s1, t1 = OT.OTpdf((f1, fx)), OT.OTpdf((g1, gx))
    w = OT.wasser(s1, t1, 'W2', returnplan=True, derivatives=True)
    out.update(plan_f=f1, plan_g=g1, plan_fx=fx, plan_gx=gx, plan_W2=w[0], plan_dW2=w[1], plan_dpos=w[2],
               plan_H=w[3], plan_dH=w[4])
    w2 = OT.wasser(s1, t1, 'W1', returnplan=True)
    out.update(plan_W1=w2[0], plan_H_nod=w2[1])
    np.savez_compressed(os.path.join(HERE, "sliced_plan.npz"), f=f, g=g, pos=pos, **out)


if __name__ == "__main__" and len(sys.argv) == 1:
    main()
    sliced_avgplan_pending = True


def sliced_avgplan():
    """(8) slice-averaged transport plans of SlicedWasserstein (libs/OTlib.py:1222-1262,1287-1318): returnplan and
    calcWplan, with and without derivatives, on a 4 x 5 grid (the derivative of the plan is n x n x n).
    Run on its own:  python tests/golden/make_golden.py avgplan"""
    rng = np.random.default_rng(2718)
    nx, ny = 4, 5
    X, Y = np.meshgrid(np.linspace(0, 1, ny), np.linspace(0, 1, nx))
    pos = np.stack([X, Y], axis=-1)
    f = rng.random((nx, ny)) + 0.05
    g = rng.random((nx, ny)) + 0.05
    out = dict(f=f, g=g, pos=pos, Nproj=5)
    mk = lambda: (OT.OTpdf((f, pos)), OT.OTpdf((g, pos)))
    for d in ("W1", "W2"):
        s, t = mk()
        r = OT.SlicedWasserstein(s, t, 5, distfunc=d, returnplan=True)
        out["rp_%s_w" % d], out["rp_%s_H" % d] = r[0], r[1]
        s, t = mk()
        r = OT.SlicedWasserstein(s, t, 5, distfunc=d, returnplan=True, derivatives=True)
        out["rpd_%s_w" % d], out["rpd_%s_dw" % d], out["rpd_%s_H" % d], out["rpd_%s_dH" % d] = r[0], r[1], r[2], r[3]
        s, t = mk()
        r = OT.SlicedWasserstein(s, t, 5, distfunc=d, calcWplan=True)
        out["cw_%s_wplan" % d], out["cw_%s_w" % d] = r[0], r[1]
        s, t = mk()
        r = OT.SlicedWasserstein(s, t, 5, distfunc=d, calcWplan=True, derivatives=True, returnplan=True)
        (out["cwd_%s_wplan" % d], out["cwd_%s_dwplan" % d], out["cwd_%s_w" % d], out["cwd_%s_dw" % d],
         out["cwd_%s_H" % d], out["cwd_%s_dH" % d]) = r
        s, t = mk()
        r = OT.SlicedWasserstein(s, t, 5, distfunc=d, calcWplan=True, calcAvgW=False)
        assert len(r) == 1 and r[0] == out["cw_%s_wplan" % d]
    np.savez_compressed(os.path.join(HERE, "sliced_avgplan.npz"), **out)
    print("sliced_avgplan", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__" and (len(sys.argv) == 1 or sys.argv[1] == "avgplan"):
    sliced_avgplan()


def cmt_optfunc():
    """(9) the reference's CMT misfit function, UNMODIFIED (libs/loc_cmt_util.py:186-306 `optfunc_OT` with its
    BuildOTobjfromWaveform / CalcWasserWaveform / arctan_trans / buildFingerprintwindows, :430-587), over the unmodified
    FingerprintLib / OTlib: 4 stations x 3 components x 61 samples set up as the CMT notebook does
    (oracle/cmt_scenario.py); pyprop8 is absent from the image, oracle/pyprop8_stub.py stands in for it.
    Run on its own:  python tests/golden/make_golden.py cmt"""
    import importlib
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import cmt_scenario, pyprop8_stub
    pyprop8_stub.install()
    cmt = importlib.import_module("libs.loc_cmt_util")
    out = {}
    for c in (False, True):
        tag = "cmt" if c else "loc"
        optdata, t = cmt_scenario.build_optdata(cmt, cmt=c)
        models = cmt_scenario.trial_models(cmt, cmt=c)
        out[tag + "_models"] = np.array(models)
        out[tag + "_obs"] = optdata["prop8data"]["obs_seis"]
        out[tag + "_grids"] = np.array(optdata["OTdata"]["obs_grids"], dtype=np.float64)
        mis, dmis, seis, jac, drs = [], [], [], [], []
        for m in models:
            a, b, tt, s = cmt.optfunc_OT(m, optdata, returnseis=True)                 # Wopt = 'Wavg' (notebook cell 34)
            mis.append(a); dmis.append(b); seis.append(s)
            a2, b2, dxyz, dr = cmt.optfunc_OT(m, optdata, returnderiv=True)            # + d(seis)/d(model), dW/d(seis)
            assert a2 == a
            jac.append(dxyz.reshape(dxyz.shape[0], -1)); drs.append(dr)
        out[tag + "_mis"], out[tag + "_dmis"], out[tag + "_seis"] = np.array(mis), np.array(dmis), np.array(seis)
        out[tag + "_J"], out[tag + "_dr"] = np.array(jac), np.array(drs)               # (M, P, nr*nc*nt), (M, nr, nc, nt)
        a, b = cmt.optfunc_OT(models[0], optdata, return2W=True)                       # both marginals
        out[tag + "_mis2W"], out[tag + "_dmis2W"] = np.array(a), np.array(b)
        for w in ("Wt", "Wu"):
            optdata["OTdata"]["Wopt"] = w
            a, b = cmt.optfunc_OT(models[1], optdata)
            out[tag + "_mis" + w], out[tag + "_dmis" + w] = np.array(a), np.array(b)
        optdata["OTdata"]["Wopt"] = "Wavg"
    np.savez_compressed(os.path.join(HERE, "cmt_optfunc.npz"), **out)
    print("cmt_optfunc", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__" and (len(sys.argv) == 1 or sys.argv[1] == "cmt"):
    cmt_optfunc()


def fd_checkers():
    """(10) the finite-difference checkers the derivative notebook calls directly on the two modules:
    fp.check_FDderiv (libs/FingerprintLib.py:516-572, Ricker_waveform_derivatives.ipynb cell 31) and
    OT._checkderivMarg (libs/OTlib.py:330-393, cell 36), on the notebook's noiseless Ricker pair.
    Run on its own:  python tests/golden/make_golden.py fd"""
    trange = [-2.0, 2.0]
    tp, wp = ru.rickerwavelet(5.0, 3.0, 0.5, trange=trange)                       # cell 7
    to, wo = ru.rickerwavelet(0.0, 1.6, 1.0, trange=trange)
    grid, lam = (-2.0, 2.0, -2.0, 3.5, 80, 512), 0.03                             # cell 12
    wfo, tgt = ru.BuildOTobjfromWaveform(to, wo, grid, lambdav=lam)
    wfp, src = ru.BuildOTobjfromWaveform(tp, wp, grid, lambdav=lam, deriv=True)
    ks = np.array([6302, 11236, 14513, 22195, 30191, 34237, 39605])               # grid points of the notebook's table
    fd = np.array([fp.check_FDderiv(wfp, int(k)) for k in ks])                    # (segment, d/du_i, d/du_i+1)
    marg = np.array([OT._checkderivMarg(src, tgt, 0.5, distfunc='W2', percent=True, ind=[int(k)], returnmargW=True)
                     for k in ks], dtype=np.float64)
    avg = np.array([OT._checkderivMarg(src, tgt, 0.5, distfunc='W2', percent=True, ind=[int(k)]) for k in ks[:3]],
                   dtype=np.float64)
    np.savez(os.path.join(HERE, "fd_checkers.npz"), tp=tp, wp=wp, to=to, wo=wo, grid=np.array(grid, dtype=np.float64),
             lam=lam, ks=ks, fd=fd, dddy=wfp.dddy[ks], marg=marg, avg=avg)
    print("fd_checkers", fd, marg, avg, sep="\n")


if __name__ == "__main__" and (len(sys.argv) == 1 or sys.argv[1] == "fd"):
    fd_checkers()


def inversion_notebook():
    """(11) Ricker_Figs_3_8.ipynb run as it is on the unmodified reference (tests/test_gpu_dropin._run_notebook: every
    code cell but the ones that draw): the L-BFGS-B inversion of cell 32.  The observed waveform carries Gaussian-process
    noise whose draw depends on the installed scikit-learn; the fixture stores it so that the GPU test knows whether it
    is looking at the same problem.  Run on its own:  python tests/golden/make_golden.py inversion"""
    sys.path.insert(0, os.path.dirname(HERE))
    from test_gpu_dropin import _run_notebook
    ns = {}
    _run_notebook("Ricker_Figs_3_8.ipynb", ns)
    o = ns["opt1"]
    np.savez(os.path.join(HERE, "inversion_notebook.npz"), wobs=ns["wobs"], x=o.x, fun=o.fun, nfev=o.nfev, nit=o.nit,
             its=np.array(ns["ricker_util_opt"].Wits), was=np.array(ns["was"], dtype=np.float64))
    print("inversion_notebook", o.x, o.fun, o.nfev, o.nit)


if __name__ == "__main__" and (len(sys.argv) == 1 or sys.argv[1] == "inversion"):
    inversion_notebook()
