"""Drop-in for the hot-path part of the reference's ``libs/FingerprintLib.py``.

Same class name, constructor signature, method names and attribute protocol as
``libs.FingerprintLib.waveformFP`` (libs/FingerprintLib.py:48-385); the
arithmetic runs in libwfot.so (include/wfot.h) on the GPU.  Heavy per-pixel
attributes (``dfield, pdf, irays, lrays, xrays, dddy, pos``) live on the device
and are copied to host NumPy arrays the first time they are read, so objects
can be stored in the adapters' history lists and pickled like the reference's.

Out of scope (raise NotImplementedError): method='FMM' / 'NNsearch'
(libs/FingerprintLib.py:139-152,160-164; marked obsolete by the reference,
unused by every notebook) and the plotting helpers.  The finite-difference checker the derivative notebook calls
(check_FDderiv, libs/FingerprintLib.py:516-572) is provided on top of the GPU distance field.
"""
from __future__ import annotations

import time

import numpy as np

from . import batch as _B


class Error(Exception):
    """Base class for other exceptions"""


class WaveformPFderivError(Exception):
    """Raised when WaveformFP.wdistderiv is called without first calling WaveformFP.wdist"""

    def __init__(self, msg=''):
        super().__init__('\n WaveformFP.wdistderiv may only be called after WaveformFP.wdist \n')


class FingerprintMethodError(Exception):
    """Raised when WaveformFP.calcpdf is called with invalid method string"""

    def __init__(self, msg=''):
        super().__init__('\n Method not recognized by WaveformFP.calcpdf \n')


class FMMlibraryError(Exception):
    """Raised when FMM library is not installed"""

    def __init__(self, msg=''):
        super().__init__('\n scikit-fmm library is not installed. see https://pypi.org/project/scikit-fmm/\n')


_LAZY = {"dfield": "dfield", "pdf": "pdf", "irays": "iray", "lrays": "lray", "xrays": "xray", "dddy": "dddy"}


class waveformFP(object):
    """Waveform object for fingerprint calculations; interface of
    libs/FingerprintLib.py:48-385."""

    def __init__(self, t, w, grid, fpgrid=None, theta=45.0, tantheta=1.0):
        (t0, t1, u0, u1, nug, ntg) = grid
        if tantheta != 1.0:                                     # libs/FingerprintLib.py:77-82
            theta = np.arctan(tantheta) * 180. / np.pi
        elif theta != 45.0:
            tantheta = np.tan(np.pi * theta / 180.0)
        else:
            tantheta = 1.0
        t = np.asarray(t)
        w = np.asarray(w)
        self.ntg = int(ntg)
        self.nug = int(nug)
        self.ulim = (u0, u1)
        self.tlim = (t0, t1)
        self.tant = tantheta
        self.theta = theta
        Delt = self.tant * (t1 - t0)
        self.tlimn = ((t[0] - t0) / Delt, (t[-1] - t0) / Delt)   # :91
        self.ulimn = (0., 1.)
        self.nt = len(t)
        if fpgrid is None:                                      # :95-100
            self.tlimfp = self.tlim
            self.ulimfp = self.ulim
            self.tlimnfp = self.tlimn
            self.ulimnfp = self.ulimn
        else:                                                   # :101-106
            (fp_t0, fp_t1, fp_u0, fp_u1) = fpgrid[0:4]
            self.tlimfp = (fp_t0, fp_t1)
            self.ulimfp = (fp_u0, fp_u1)
            self.tlimnfp = ((fp_t0 - t0) / Delt, (fp_t1 - t0) / Delt)
            self.ulimnfp = ((fp_u0 - u0) / (u1 - u0), (fp_u1 - u0) / (u1 - u0))
        self.delgrid = np.array([(self.ulimnfp[1] - self.ulimnfp[0]) / self.nug,
                                 (self.tlimnfp[1] - self.tlimnfp[0]) / self.ntg])
        self.p = np.array([t, w]).T                              # :109
        self._fpgrid = None if fpgrid is None else tuple(fpgrid[0:4])
        self._grid = (t0, t1, u0, u1, self.nug, self.ntg)
        self._dev = {}       # device tensors of the last calcpdf()
        self._host = {}      # host copies materialised on attribute access
        self._geom = None
        self.dcalc = False
        self.drcalc = False

    # --- geometry attributes (pn, x0, delta_n, lsq_n): tiny, computed by the kernel's prep
    def _geometry(self):
        if self._geom is None:
            if "pn" in self._dev:
                pn = self._dev["pn"][0].cpu().numpy()
            else:   # before calcpdf(): host arithmetic identical to libs/FingerprintLib.py:110
                t0, t1, u0, u1 = self._grid[:4]
                Delt = self.tant * (t1 - t0)
                pn = np.array([(self.p.T[0] - t0) / Delt, (self.p.T[1] - u0) / (u1 - u0)]).T
            delta = np.subtract(pn[1:], pn[:-1])
            self._geom = dict(pn=pn, x0=pn[:-1].reshape(1, self.nt - 1, 2), delta_n=delta,
                              lsq_n=np.sum(np.multiply(delta, delta), axis=1))
        return self._geom

    pn = property(lambda self: self._geometry()["pn"])
    x0 = property(lambda self: self._geometry()["x0"])
    delta_n = property(lambda self: self._geometry()["delta_n"])
    lsq_n = property(lambda self: self._geometry()["lsq_n"])

    def calcpdf(self, q=None, lambdav=0.04, deriv=False, method='Enumerate', verbose=False, nsegs=0):
        """libs/FingerprintLib.py:117-180 (method 'Enumerate')."""
        self.lam = lambdav
        if method in ('FMM', 'fmm', 'NNsearch'):
            raise NotImplementedError("waveform_ot_b200: method=%r is out of scope (reference marks it "
                                      "obsolete); only 'Enumerate' is implemented" % method)
        if method != 'Enumerate':
            print(' Method string provided = ' + method)
            raise FingerprintMethodError
        if q is not None and q != 2:
            raise NotImplementedError("q must be None or 2 (the reference leaves .pdf unset otherwise)")
        t0 = time.time()
        self.wdist(deriv=deriv, _q=q)
        self.type = 'Enu'
        self.tcalc_fp = time.time() - t0
        self.q = q
        self.tcalc_pdf = 0.0          # the density is produced by the same kernel as the distance field
        if verbose:
            print(' calcpdf:\n' + ' Time taken for distance field:', self.tcalc_fp,
                  '\n Time taken for pdf field:', self.tcalc_pdf)

    def wdist(self, deriv=False, _q="keep"):
        """libs/FingerprintLib.py:230-272: distance field, nearest segment, (optionally) derivatives."""
        import torch
        q = getattr(self, "q", None) if _q == "keep" else _q
        lam = getattr(self, "lam", 0.04)
        out = _B.fingerprint_batch(self.p.T[0], self.p.T[1], self._grid, self.nug, self.ntg, lam, q=q,
                                   tantheta=self.tant, fpgrids=self._fpgrid, deriv=deriv)
        torch.cuda.current_stream().synchronize()
        self._dev = {k: v for k, v in out.items() if k in ("pn", "dfield", "iray", "lray", "xray", "pdf", "dddy")}
        self._host = {}
        self._geom = None
        self._status = out["status"].raise_for_reference(derivatives=deriv, what="waveformFP.wdist")
        self.dcalc = True
        if deriv:
            self.drcalc = True

    def wdistderiv(self, verbose=False):
        """libs/FingerprintLib.py:333-385."""
        if not self.dcalc:
            raise WaveformPFderivError
        if "dddy" not in self._dev:
            self.wdist(deriv=True)

    def __getattr__(self, name):
        # lazily materialised per-pixel fields (only called when normal lookup fails)
        if name in _LAZY:
            d = self.__dict__
            key = _LAZY[name]
            if key in d.get("_host", {}):
                return d["_host"][key]
            if key in d.get("_dev", {}):
                v = d["_dev"][key][0].cpu().numpy()
                if name == "irays":
                    v = v.astype(np.int64)
                d["_host"][key] = v
                return v
        if name == "pos" and self.__dict__.get("dcalc"):
            Xn, Yn = np.meshgrid(np.linspace(self.tlimnfp[0], self.tlimnfp[1], self.ntg),
                                 np.linspace(self.ulimnfp[0], self.ulimnfp[1], self.nug))   # :254,269
            pos = np.dstack((Xn, Yn))
            self.__dict__.setdefault("_host", {})["pos"] = pos
            self.__dict__["pos"] = pos
            return pos
        raise AttributeError(name)

    def _chain_deriv(self, chains):
        import torch
        need = ("pdf", "dfield", "iray", "dddy")
        if any(k not in self._dev for k in need):
            raise WaveformPFderivError
        out = _B.pdfderiv_batch(self._dev["pdf"], self._dev["dfield"], self._dev["iray"], self._dev["dddy"],
                                chains, self.nt, self.lam, q=self.q)
        torch.cuda.current_stream().synchronize()
        return out[0].cpu().numpy()

    def PDFderiv(self, chainmatrix=None):
        """libs/FingerprintLib.py:182-203."""
        ch = None
        if type(chainmatrix) == np.ndarray:
            ch = chainmatrix.reshape(1, 1, -1)
        self.pdfd = self._chain_deriv(ch)[0]

    def PDFderivMarg(self, chainmatrix):
        """libs/FingerprintLib.py:205-228."""
        ch = np.stack([np.asarray(chainmatrix[0]).reshape(-1), np.asarray(chainmatrix[1]).reshape(-1)])[None]
        r = self._chain_deriv(ch)
        self.pdfdMarg = [r[0], r[1]]

    # --- pickling: host arrays only (no device handles), like the reference's objects
    def __getstate__(self):
        st = dict(self.__dict__)
        host = dict(st.get("_host", {}))
        for name, key in _LAZY.items():
            if key in st.get("_dev", {}) and key not in host:
                v = st["_dev"][key][0].cpu().numpy()
                host[key] = v.astype(np.int64) if name == "irays" else v
        st["_host"] = host
        st["_dev"] = {}
        geom = self._geometry()
        st["_geom"] = geom
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)


def check_FDderiv(wf, k, du=0.001, verbose=False):
    """libs/FingerprintLib.py:516-572 (Ricker_waveform_derivatives.ipynb cell 31): central finite differences of the
    distance at grid point k with respect to the amplitudes of the two samples that bound its nearest segment.
    Returns (segment, d d[k]/d u_segment, d d[k]/d u_segment+1).  Each of the four perturbed windows is one call of
    the distance-field kernel (the reference evaluates all grid points too, wavedistv :456-474, and reads point k);
    like the reference, the perturbed objects are built on (tlim, ulim) without the fpgrid of `wf`."""
    t, RF = wf.p.T[0], wf.p.T[1]
    u0, u1 = wf.ulim
    t0, t1 = wf.tlim
    i = int(wf.irays[k])
    dups = du * np.abs(RF[i])                                        # :527

    def dist_at_k(j, sign):
        RFq = np.copy(RF)
        RFq[j] += sign * dups
        wq = waveformFP(t, RFq, (t0, t1, u0, u1, wf.nug, wf.ntg), tantheta=wf.tant)
        wq.wdist()
        return wq.dfield.reshape(-1)[k], int(wq.irays[k])

    (dp0, ip0), (dm0, im0) = dist_at_k(i, +1.0), dist_at_k(i, -1.0)
    (dp1, ip1), (dm1, im1) = dist_at_k(i + 1, +1.0), dist_at_k(i + 1, -1.0)
    dddy0fd = (dp0 - dm0) / (2 * dups)                               # :544
    dddy1fd = (dp1 - dm1) / (2 * dups)                               # :560
    if verbose:
        print('\n segments after FD perturbation : ', ' pos 0 ', ip0, ' minus 0', im0, 'pos 1 ', ip1, ' minus 1', im1)
    return i, dddy0fd, dddy1fd


# Names of libs/FingerprintLib.py that are outside the hot path (plotting, the obsolete FMM / nearest-neighbour methods,
# point-wise host evaluators): with the shim installed over the reference package they are served by
# the reference's own module (adapters.reference_attr); otherwise asking for one says so instead of a bare AttributeError.
_OUT_OF_SCOPE = ("NNsearch", "wavedist", "wavedistv", "wavederiv", "check_FDchain", "wPDFderiv", "plot_RF_SDF",
                 "plotPDFsurface", "plot_phi", "plot_rays_discrete", "plot_rays", "plot_LS", "plot_2LS", "plotMarginals",
                 "calcFMM_dist_deriv", "find_raystart_point_with_gradient")


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        try:   # installed over the reference package (adapters.install): the reference's own function serves the call
            from . import adapters
            return adapters.reference_attr("FingerprintLib", name)
        except AttributeError:
            pass
        raise AttributeError("waveform_ot_b200.FingerprintLib: %r of libs/FingerprintLib.py is outside the accelerated "
                             "path (plotting, FMM / NNsearch, host-side point evaluators); use the reference module "
                             "for it" % name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
