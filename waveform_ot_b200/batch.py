"""Batched drivers over the C ABI (include/wfot.h).

These are the entry points the reference's per-window Python loops map onto
(libs/loc_cmt_util.py:256-271, 503-519 loop over stations x components;
notebook sweeps loop over trial models).  torch is used only to own device
memory and streams; every computation happens inside libwfot.so.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _cabi as C

GRID_DTYPE = np.dtype([("t0", "f8"), ("t1", "f8"), ("u0", "f8"), ("u1", "f8"),
                       ("fp_t0", "f8"), ("fp_t1", "f8"), ("fp_u0", "f8"), ("fp_u1", "f8"),
                       ("tantheta", "f8"), ("has_fpgrid", "i4"), ("reserved", "i4")])
assert GRID_DTYPE.itemsize == ctypes.sizeof(C.wfot_grid)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)   # the handle without building a Stream object
_cuda_checked = False


def _stream():
    """cudaStream_t of torch's current stream on the current device (what every C-ABI call is queued on)."""
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _device():
    global _cuda_checked
    if not _cuda_checked:      # asked once per process: torch.cuda.is_available() costs microseconds on every call
        if not torch.cuda.is_available():
            raise RuntimeError("waveform_ot_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _cuda_checked = True
    return torch.device("cuda", torch.cuda.current_device())


def pack_grids(grids, tantheta=1.0, fpgrids=None):
    """grids: one (t0,t1,u0,u1,...) tuple or a sequence of them -> (n, 80-byte) device tensor."""
    if np.isscalar(grids[0]):
        grids = [grids]
    n = len(grids)
    tan = np.broadcast_to(np.asarray(tantheta, dtype=np.float64), (n,))
    arr = np.zeros(n, dtype=GRID_DTYPE)
    for i, g in enumerate(grids):
        arr[i]["t0"], arr[i]["t1"], arr[i]["u0"], arr[i]["u1"] = g[0], g[1], g[2], g[3]
        arr[i]["tantheta"] = tan[i]
        fg = None if fpgrids is None else (fpgrids if np.isscalar(fpgrids[0]) else fpgrids[i])
        if fg is not None:
            arr[i]["fp_t0"], arr[i]["fp_t1"], arr[i]["fp_u0"], arr[i]["fp_u1"] = fg[0], fg[1], fg[2], fg[3]
            arr[i]["has_fpgrid"] = 1
    return torch.from_numpy(arr.view(np.uint8).reshape(n, GRID_DTYPE.itemsize).copy()).to(_device())


def _as_device(x, dtype=None):
    """numpy / torch (host or device) -> contiguous device tensor; float32/float64 kept."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.ascontiguousarray(x)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        t = torch.from_numpy(a)
    if dtype is not None:
        t = t.to(dtype)
    return t.to(_device(), non_blocking=True).contiguous()


def _dt(t):
    if t.dtype == torch.float32:
        return C.F32
    if t.dtype == torch.float64:
        return C.F64
    raise TypeError("waveform arrays must be float32 or float64")


class Status:
    """Per-call data-dependent condition counters (WFOT_STAT_* slots)."""

    def __init__(self):
        self.t = torch.zeros(C.STAT_SLOTS, dtype=torch.int32, device=_device())

    def read(self):
        return self.t.cpu().numpy()

    def raise_for_reference(self, derivatives=True, what="waveform_ot_b200"):
        """Map the counters onto what the reference does on the same data (call after the stream has been
        synchronised; costs one 32-byte device -> host copy):
          * common_cdf > 0  -> OTlib.TargetSourceCDFError: MargWasserstein always calls
            wasser(checkCommonCDF=True) (libs/OTlib.py:1111-1113, 663-666) and the derivatives are
            invalid when source and target CDFs share values;
          * neg_pdf > 0     -> OTlib.PDFSignError (libs/OTlib.py:91);
          * zero_dist > 0 with derivatives -> RuntimeWarning: a pixel lies exactly on the waveform, d = 0,
            and d(d)/dw = 0/0 = NaN there exactly as in libs/FingerprintLib.py:355 (NumPy prints the same
            kind of warning); the NaN is in the returned gradient of the two samples of that segment;
          * degenerate > 0  -> RuntimeWarning: zero-length segments (repeated samples).  The reference
            divides 0/0 at libs/FingerprintLib.py:257 and returns NaN fields; here such a segment acts as
            a point.
        Returns the counters as a NumPy array."""
        import warnings
        st = self.read()
        if st[C.STAT_NEG_PDF] or st[C.STAT_COMMON_CDF]:
            from . import OTlib
            if st[C.STAT_NEG_PDF]:
                raise OTlib.PDFSignError()
            raise OTlib.TargetSourceCDFError(["%d common value(s) in %s" % (int(st[C.STAT_COMMON_CDF]), what)])
        if derivatives and st[C.STAT_ZERO_DIST]:
            warnings.warn("%s: %d pixel(s) at zero distance from the waveform: d(d)/dw is NaN there "
                          "(libs/FingerprintLib.py:355)" % (what, int(st[C.STAT_ZERO_DIST])), RuntimeWarning, stacklevel=3)
        if st[C.STAT_DEGENERATE_SEG]:
            warnings.warn("%s: %d zero-length waveform segment(s) (repeated samples); the reference returns NaN "
                          "fields for such input (libs/FingerprintLib.py:257)" % (what, int(st[C.STAT_DEGENERATE_SEG])),
                          RuntimeWarning, stacklevel=3)
        return st

    def scan_pairs(self):
        """Executed (pixel, segment) pairs of the pruned scan (64-bit counter in slots 6-7,
        kept by the kernels in units of 2048 pairs)."""
        return 2048 * int(self.t.cpu().numpy()[C.STAT_SCAN_TILES:C.STAT_SCAN_TILES + 2].view("int64")[0])


def fingerprint_batch(t, w, grids, nug, ntg, lambdav, q=None, tantheta=1.0, fpgrids=None,
                      deriv=False, fields=("dfield", "iray", "lray", "xray", "pdf", "pn"), status=None):
    """waveformFP(...).calcpdf(...) for B windows.  t: (nt,) shared or (B, nt); w: (B, nt).
    Returns dict of device tensors (FP64; iray int32)."""
    dev = _device()
    w = _as_device(w)
    if w.dim() == 1:
        w = w[None, :]
    B, nt = w.shape
    t = _as_device(t, w.dtype)
    t_stride = 0 if t.dim() == 1 else nt
    g = grids if isinstance(grids, torch.Tensor) else pack_grids(grids, tantheta, fpgrids)
    npix = nug * ntg
    out = {}
    f64 = dict(dtype=torch.float64, device=dev)
    want = set(fields) | ({"dddy"} if deriv else set())
    if "pn" in want: out["pn"] = torch.empty((B, nt, 2), **f64)
    if "dfield" in want: out["dfield"] = torch.empty((B, nug, ntg), **f64)
    if "iray" in want: out["iray"] = torch.empty((B, npix), dtype=torch.int32, device=dev)
    if "lray" in want: out["lray"] = torch.empty((B, npix), **f64)
    if "xray" in want: out["xray"] = torch.empty((B, npix, 2), **f64)
    if "pdf" in want: out["pdf"] = torch.empty((B, nug, ntg), **f64)
    if "dddy" in want: out["dddy"] = torch.empty((B, npix, 2), **f64)
    wsb = C.lib.wfot_fingerprint_workspace_bytes(B, nt, nug, ntg)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = status or Status()
    C.check(C.lib.wfot_fingerprint_batch(
        C.ptr(t), C.ptr(w), _dt(w), t_stride, nt, C.ptr(g), g.shape[0], B, nug, ntg,
        float(lambdav), 0 if q is None else int(q),
        C.ptr(out.get("pn")), C.ptr(out.get("dfield")), C.ptr(out.get("iray")), C.ptr(out.get("lray")),
        C.ptr(out.get("xray")), C.ptr(out.get("pdf")), C.ptr(out.get("dddy")),
        C.ptr(ws), wsb, C.ptr(st.t), _stream()), "wfot_fingerprint_batch")
    out["status"] = st
    out["_keepalive"] = (t, w, g, ws)
    return out


def marginals_batch(pdf, status=None):
    """OTpdf (2-D) + setMarginals for B densities (B, nug, ntg) -> amp, marg_t, marg_u."""
    dev = _device()
    pdf = _as_device(pdf, torch.float64)
    if pdf.dim() == 2:
        pdf = pdf[None]
    B, nug, ntg = pdf.shape
    amp = torch.empty(B, dtype=torch.float64, device=dev)
    mt = torch.empty((B, ntg), dtype=torch.float64, device=dev)
    mu = torch.empty((B, nug), dtype=torch.float64, device=dev)
    st = status or Status()
    C.check(C.lib.wfot_marginals_batch(C.ptr(pdf), B, nug, ntg, C.ptr(amp), C.ptr(mt), C.ptr(mu),
                                       C.ptr(st.t), _stream()), "wfot_marginals_batch")
    return dict(amp=amp, marg_t=mt, marg_u=mu, status=st)


def otpdf1d_batch(f, status=None):
    """OTpdf.__init__ for B 1-D densities (B, n): amp, normalised pdf, CDF."""
    dev = _device()
    f = _as_device(f)
    if f.dim() == 1:
        f = f[None]
    B, n = f.shape
    f64 = dict(dtype=torch.float64, device=dev)
    amp, pdfn, cdf = torch.empty(B, **f64), torch.empty((B, n), **f64), torch.empty((B, n), **f64)
    st = status or Status()
    C.check(C.lib.wfot_otpdf1d_batch(C.ptr(f), _dt(f), n, B, C.ptr(amp), C.ptr(pdfn), C.ptr(cdf),
                                     C.ptr(st.t), _stream()), "wfot_otpdf1d_batch")
    return dict(amp=amp, pdf=pdfn, cdf=cdf, status=st, _keepalive=(f,))


def ot1d_batch(f, g, xf, xg, distfunc="W12", derivatives=False, want_cdf=False, want_merge=False,
               status=None):
    """OTpdf (1-D) + wasser for B pairs.  f (B,n) or (n,); g (B,m) or (m,) shared; x likewise."""
    dev = _device()
    f = _as_device(f)
    g = _as_device(g, f.dtype)
    if f.dim() == 1:
        f = f[None]
    B, n = f.shape
    m = g.shape[-1]
    xf = _as_device(xf, torch.float64)
    xg = _as_device(xg, torch.float64)
    pmask = {"W1": C.W1, "W2": C.W2, "W12": C.W12}[distfunc]
    f64 = dict(dtype=torch.float64, device=dev)
    Wd = torch.zeros((2, B, 2), **f64)      # W and dpos side by side: one device -> host copy brings both
    W = Wd[0]
    dpos = Wd[1] if derivatives else None
    dW1 = torch.empty((B, n), **f64) if derivatives and pmask & 1 else None
    dW2 = torch.empty((B, n), **f64) if derivatives and pmask & 2 else None
    amp = torch.empty(B, **f64)
    cdf_f = torch.empty((B, n), **f64) if want_cdf else None
    cdf_g = torch.empty((B, m), **f64) if want_cdf else None
    merge = torch.empty((B, n + m - 1), dtype=torch.int32, device=dev) if want_merge else None
    st = status or Status()
    C.check(C.lib.wfot_ot1d_batch(
        C.ptr(f), C.ptr(g), _dt(f), C.ptr(xf), C.ptr(xg),
        n, 0 if g.dim() == 1 else m, 0 if xf.dim() == 1 else n, 0 if xg.dim() == 1 else m,
        n, m, B, pmask, int(bool(derivatives)),
        C.ptr(W), C.ptr(dW1), C.ptr(dW2), C.ptr(dpos), C.ptr(amp), C.ptr(cdf_f), C.ptr(cdf_g),
        C.ptr(merge), C.ptr(st.t), _stream()), "wfot_ot1d_batch")
    return dict(W=W, dW1=dW1, dW2=dW2, dpos=dpos, amp=amp, cdf_f=cdf_f, cdf_g=cdf_g, merge_order=merge,
                status=st, W_dpos=Wd, _keepalive=(f, g, xf, xg))


def plan_batch(r, n, m, perm_f=None, perm_g=None, accumulate=False, derivatives=False):
    """Transport plan(s) from an ot1d_batch(..., want_cdf=True, want_merge=True) result `r`
    (libs/OTlib.py:718-740): H (B, n, m) and, with derivatives, dH (B, n, n, m); with accumulate=True the
    pairs add into one H (n, m) / dH (n, n, m), rows / columns scattered through perm_f (B, n) / perm_g (B, m)
    (libs/OTlib.py:1247-1262)."""
    dev = _device()
    Bn = r["cdf_f"].shape[0]
    f64 = dict(dtype=torch.float64, device=dev)
    lead = () if accumulate else (Bn,)
    H = torch.empty(lead + (n, m), **f64)
    dH = torch.empty(lead + (n, n, m), **f64) if derivatives else None
    as_i32 = lambda p: None if p is None else torch.as_tensor(np.ascontiguousarray(p), dtype=torch.int32).to(dev).contiguous()
    pf, pg = as_i32(perm_f), as_i32(perm_g)
    C.check(C.lib.wfot_plan_batch(C.ptr(r["cdf_f"]), C.ptr(r["cdf_g"]), C.ptr(r["merge_order"]), C.ptr(r["amp"]),
                                  C.ptr(pf), C.ptr(pg), n, m, Bn, int(bool(accumulate)), C.ptr(H), C.ptr(dH), _stream()),
            "wfot_plan_batch")
    return H, dH, (pf, pg)


def pdfderiv_batch(pdf, dfield, iray, dddy, chain, nt, lambdav, q=None):
    """PDFderiv / PDFderivMarg for B windows.  chain: None, (B,npix) or (B,nchain,npix)."""
    dev = _device()
    pdf = _as_device(pdf, torch.float64)
    B = pdf.shape[0]
    npix = pdf[0].numel()
    dfield = _as_device(dfield, torch.float64) if dfield is not None else None
    iray = iray.to(dev).to(torch.int32).contiguous()
    dddy = _as_device(dddy, torch.float64)
    nchain = 1
    if chain is not None:
        chain = _as_device(chain, torch.float64).reshape(B, -1, npix)
        nchain = chain.shape[1]
    out = torch.empty((B, nchain, nt), dtype=torch.float64, device=dev)
    C.check(C.lib.wfot_pdfderiv_batch(C.ptr(pdf), C.ptr(dfield), C.ptr(iray), C.ptr(dddy), C.ptr(chain),
                                      nchain, B, npix, nt, float(lambdav), 0 if q is None else int(q),
                                      C.ptr(out), _stream()), "wfot_pdfderiv_batch")
    return out


class Target:
    """Observed-window marginals in the form the fused kernel consumes: CDF + bin
    positions of the time and amplitude marginals (what `wfobs_target.marg[i].cdf/.x`
    hold in the reference, libs/OTlib.py:112-114,157-160)."""

    def __init__(self, cdf_t, x_t, cdf_u, x_u):
        self.cdf_t, self.x_t, self.cdf_u, self.x_u = cdf_t, x_t, cdf_u, x_u
        self.per_window = cdf_t.dim() == 2 and cdf_t.shape[0] > 1

    @property
    def rows(self):
        """Number of observed windows held: window b of a batch is compared with row b % rows (1 = one
        observation for the whole batch, B = one per window, nr*nc = one per station/component)."""
        return int(self.cdf_t.shape[0]) if self.cdf_t.dim() == 2 else 1

    @staticmethod
    def from_waveform(t, w, grids, nug, ntg, lambdav, q=None, tantheta=1.0, fpgrids=None, transform=False,
                      status=None):
        """Observed window(s) -> marginal CDFs + bin positions (wfot_marginal_cdfs_batch: the fused path's own
        kernels and summation orders, so a predicted window equal to the observed one yields bit-identical
        CDFs and the common-CDF condition of libs/OTlib.py:663-666 is detected exactly).
        transform=True applies the arctan amplitude transform in-kernel with each grid's (u0, u1), as
        misfit_grad_batch(transform=True) does for the predicted windows."""
        dev = _device()
        w = _as_device(w)
        if w.dim() == 1:
            w = w[None, :]
        B, nt = w.shape
        t = _as_device(t, w.dtype)
        t_stride = 0 if t.dim() == 1 else nt
        g = grids if isinstance(grids, torch.Tensor) else pack_grids(grids, tantheta, fpgrids)
        f64 = dict(dtype=torch.float64, device=dev)
        ct, cu = torch.empty((B, ntg), **f64), torch.empty((B, nug), **f64)
        amp = torch.empty(B, **f64)
        wsb = C.lib.wfot_misfit_grad_workspace_bytes(B, nt, nug, ntg)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        st = status or Status()
        rc = C.lib.wfot_marginal_cdfs_batch(
            C.ptr(t), C.ptr(w), _dt(w), t_stride, nt, C.ptr(g), g.shape[0], B, nug, ntg, float(lambdav),
            0 if q is None else int(q), int(bool(transform)), C.ptr(ct), C.ptr(cu), C.ptr(amp),
            C.ptr(ws), wsb, C.ptr(st.t), _stream())
        if rc == C.ERR_UNSUPPORTED and not transform:
            # window too long for the fused kernels' shared memory (they cannot evaluate a predicted window
            # against this target either): the materialising kernels still give the CDFs
            fp = fingerprint_batch(t, w, g, nug, ntg, lambdav, q=q, fields=("pdf",), status=st)
            mg = marginals_batch(fp["pdf"], status=st)
            ct = otpdf1d_batch(mg["marg_t"], status=st)["cdf"]
            cu = otpdf1d_batch(mg["marg_u"], status=st)["cdf"]
            amp = mg["amp"]
        else:
            C.check(rc, "wfot_marginal_cdfs_batch")
        # bin positions = pixel axes (libs/OTlib.py:157-158 on wf.pos): host arithmetic, IEEE-identical to the
        # kernel's (t - t0) / (tan(theta) (t1 - t0)) and np.linspace
        ends = t[..., [0, -1]].to(torch.float64).cpu().numpy().reshape(-1, 2)
        gr = g.cpu().numpy().view(GRID_DTYPE).reshape(-1)
        xt, xu = np.empty((B, ntg)), np.empty((B, nug))
        for b in range(B):
            gb = gr[b % len(gr)]
            e = ends[b if t_stride else 0]
            delt = gb["tantheta"] * (gb["t1"] - gb["t0"])
            u0, du = (0.0, 1.0) if transform else (gb["u0"], gb["u1"] - gb["u0"])
            if gb["has_fpgrid"]:
                a0, a1 = (gb["fp_t0"] - gb["t0"]) / delt, (gb["fp_t1"] - gb["t0"]) / delt
                c0, c1 = (gb["fp_u0"] - u0) / du, (gb["fp_u1"] - u0) / du
            else:
                a0, a1, c0, c1 = (e[0] - gb["t0"]) / delt, (e[1] - gb["t0"]) / delt, 0.0, 1.0
            xt[b] = np.linspace(a0, a1, ntg)
            xu[b] = np.linspace(c0, c1, nug)
        tg = Target(ct, torch.from_numpy(xt).to(dev), cu, torch.from_numpy(xu).to(dev))
        tg.per_window = B > 1
        tg.amp = amp
        tg.status = st
        tg._keepalive = (t, w, g, ws)
        return tg


def pixel_axes(pn, grids_dev, nug, ntg):
    """FP64 pixel axes exactly as np.linspace builds them (libs/FingerprintLib.py:254):
    i*step + start, last element = stop.  pn: (B, nt, 2) device tensor."""
    B = pn.shape[0]
    gr = grids_dev.cpu().numpy().view(GRID_DTYPE).reshape(-1)
    pnh = pn[:, [0, -1], 0].cpu().numpy()
    xt = np.empty((B, ntg))
    xu = np.empty((B, nug))
    for b in range(B):
        g = gr[0 if len(gr) == 1 else b]
        if g["has_fpgrid"]:
            delt = g["tantheta"] * (g["t1"] - g["t0"])
            a0, a1 = (g["fp_t0"] - g["t0"]) / delt, (g["fp_t1"] - g["t0"]) / delt
            du = g["u1"] - g["u0"]
            c0, c1 = (g["fp_u0"] - g["u0"]) / du, (g["fp_u1"] - g["u0"]) / du
        else:
            a0, a1, c0, c1 = pnh[b, 0], pnh[b, 1], 0.0, 1.0
        xt[b] = np.linspace(a0, a1, ntg)
        xu[b] = np.linspace(c0, c1, nug)
    return torch.from_numpy(xt).to(pn.device), torch.from_numpy(xu).to(pn.device)


def misfit_grad_batch(t, w, grids, nug, ntg, lambdav, target: Target, distfunc="W2", q=None,
                      tantheta=1.0, fpgrids=None, transform=False, want_grad=True, status=None,
                      workspace=None, out=None):
    """Fused evaluation of B windows: returns W (B,2) [W^t, W^u], grad (B,2,nt), dwg (B,)
    (dW^t/d(translation) in normalised time units).  Inputs may already be device tensors.
    distfunc "W12" (misfit only, want_grad=False): both orders from one fingerprint, W (B,4) =
    [W1^t, W1^u, W2^t, W2^u], dwg (B,2).
    `workspace` / `out` (a previous result dict) let a caller that streams batches of one shape
    through the library re-use the scratch and result buffers instead of allocating per call."""
    dev = _device()
    w = _as_device(w)
    if w.dim() == 1:
        w = w[None, :]
    B, nt = w.shape
    t = _as_device(t, w.dtype)
    t_stride = 0 if t.dim() == 1 else nt
    g = grids if isinstance(grids, torch.Tensor) else pack_grids(grids, tantheta, fpgrids)
    pmask = {"W1": C.W1, "W2": C.W2, "W12": C.W12}[distfunc]
    if pmask == C.W12 and want_grad:
        raise ValueError("distfunc 'W12' is a misfit-only mode of the fused path (want_grad=False); "
                         "MargWasserstein itself rejects 'W12' (libs/OTlib.py:1090-1091)")
    nW = 4 if pmask == C.W12 else 2
    f64 = dict(dtype=torch.float64, device=dev)
    if out is not None and out["W"].shape == (B, nW) and (not want_grad or (out.get("grad") is not None
                                                                           and out["grad"].shape == (B, 2, nt))):
        W, grad, dwg = out["W"], (out["grad"] if want_grad else None), out["dwg"]
    else:
        W = torch.empty((B, nW), **f64)
        grad = torch.empty((B, 2, nt), **f64) if want_grad else None
        dwg = torch.empty((B, 2) if pmask == C.W12 else B, **f64)
    wsb = C.lib.wfot_misfit_grad_workspace_bytes(B, nt, nug, ntg)
    ws = workspace if workspace is not None and workspace.numel() >= wsb else \
        torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = status or Status()
    C.check(C.lib.wfot_misfit_grad_batch(
        C.ptr(t), C.ptr(w), _dt(w), t_stride, nt, C.ptr(g), g.shape[0], B, nug, ntg,
        float(lambdav), 0 if q is None else int(q), pmask, int(bool(transform)),
        C.ptr(target.cdf_t), C.ptr(target.x_t), C.ptr(target.cdf_u), C.ptr(target.x_u),
        target.rows, C.ptr(W), C.ptr(grad), C.ptr(dwg), C.ptr(ws), ws.numel(),
        C.ptr(st.t), _stream()), "wfot_misfit_grad_batch")
    return dict(W=W, grad=grad, dwg=dwg, status=st, _keepalive=(t, w, g, ws))


def chain_batch(J, dr):
    """g_m = J_m . dr_m  (libs/ricker_util.py:399-400, libs/loc_cmt_util.py:283-296).
    J (M,P,L) or (P,L) shared; dr (M,L)."""
    dev = _device()
    J = _as_device(J, torch.float64)
    dr = _as_device(dr, torch.float64)
    if dr.dim() == 1:
        dr = dr[None]
    M, L = dr.shape
    P = J.shape[-2]
    out = torch.empty((M, P), dtype=torch.float64, device=dev)
    C.check(C.lib.wfot_chain_batch(C.ptr(J), C.ptr(dr), P, L, M, 0 if J.dim() == 2 else P * L,
                                   C.ptr(out), _stream()), "wfot_chain_batch")
    return out


def ricker_batch(params, trange=(-2.0, 2.0), deriv=False):
    """rickerwavelet(tpert, amp, f, trange, deriv) for M parameter rows (libs/ricker_util.py:38-89,
    noise free).  Returns device tensors t (M,256), w (M,256) and, with deriv, dw (M,3,256)."""
    dev = _device()
    params = _as_device(params, torch.float64).reshape(-1, 3).contiguous()
    M = params.shape[0]
    f64 = dict(dtype=torch.float64, device=dev)
    t = torch.empty((M, 256), **f64)
    w = torch.empty((M, 256), **f64)
    dw = torch.empty((M, 3, 256), **f64) if deriv else None
    C.check(C.lib.wfot_ricker_batch(C.ptr(params), M, float(trange[0]), float(trange[1]), C.ptr(t), C.ptr(w),
                                    C.ptr(dw), _stream()), "wfot_ricker_batch")
    return dict(t=t, w=w, dw=dw, _keepalive=(params,))


def sum_windows(x, out=None, workspace=None):
    """Deterministic FP64 sum over the leading (window) axis of a device tensor (B, ...)."""
    dev = _device()
    x = x.contiguous()
    B = x.shape[0]
    Cn = x[0].numel()
    if out is None:
        out = torch.empty(x.shape[1:], dtype=torch.float64, device=dev)
    wsb = C.lib.wfot_sum_windows_workspace_bytes(Cn)
    ws = workspace if workspace is not None and workspace.numel() >= wsb else \
        torch.empty(wsb, dtype=torch.uint8, device=dev)
    C.check(C.lib.wfot_sum_windows(C.ptr(x), B, Cn, C.ptr(out), C.ptr(ws), wsb, _stream()), "wfot_sum_windows")
    return out
