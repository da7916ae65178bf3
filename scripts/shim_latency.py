"""Latency of the reference's own call sequence through the drop-in shim (cfg1 shape, one evaluation):
waveformFP -> calcpdf(deriv) -> OTpdf -> MargWasserstein(derivatives, returnmargW) -> PDFderivMarg
(libs/ricker_util.py:250-268,321-337).  The unmodified reference needs 0.73 s for this on one core."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from waveform_ot_b200 import FingerprintLib as fp, OTlib as OT

f = 1.0 * 25 * 4 / 128
tt = np.arange(-2.0, (4 - 4 / 128) / 2, 4 / 128)
w0 = (1.0 - 2.0 * np.pi ** 2 * f ** 2 * tt ** 2) * np.exp(-np.pi ** 2 * f ** 2 * tt ** 2)
t = np.linspace(-2.0, 2.0, 256)
wo = 1.6 * np.concatenate((w0, w0))
grid = (-2, 2, -1.8, 4.2, 80, 512)

def build(tshift, amp):
    wf = fp.waveformFP(t + tshift, amp * np.concatenate((w0, w0)), grid)
    wf.calcpdf(lambdav=0.03, deriv=True)
    return wf, OT.OTpdf((wf.pdf, wf.pos))

wfo, tgt = build(0.0, 1.6)
def one():
    wf, src = build(0.7, 1.3)
    W, dW, dwg = OT.MargWasserstein(src, tgt, distfunc="W2", derivatives=True, returnmargW=True)
    wf.PDFderivMarg(dW)
    return W, wf.pdfdMarg
one()
t0 = time.perf_counter()
for _ in range(10):
    W, g = one()
print("shim call sequence, cfg1 shape: %.2f ms per evaluation; W = %s" % ((time.perf_counter() - t0) / 10 * 1e3, W))
