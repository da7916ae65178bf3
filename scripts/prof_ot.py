"""Timing / ncu driver for the batched 1-D OT kernel (cfg2 shape). usage: prof_ot.py <pairs> [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from waveform_ot_b200 import batch as B

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = 1024
f = torch.rand(nb, n, device="cuda") + 1e-3
g = torch.rand(nb, n, device="cuda") + 1e-3
x = torch.linspace(0, 1, n, dtype=torch.float64, device="cuda")
B.ot1d_batch(f, g, x, x, "W12", derivatives=True)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(reps):
    s.record(); r = B.ot1d_batch(f, g, x, x, "W12", derivatives=True); e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    print("ot1d W12+deriv: B=%d %.3f ms  %.2f Mpairs/s  %.1f GB/s algorithmic (12 KiB/pair)  %.1f GB/s moved (24 KiB/pair)" % (
        nb, ms, nb / ms / 1e3, nb * 12288 / ms / 1e6, nb * 24576 / ms / 1e6))
s.record(); r = B.ot1d_batch(f, g, x, x, "W12", derivatives=False); e.record(); torch.cuda.synchronize()
print("ot1d W12 only: %.3f ms  %.2f Mpairs/s" % (s.elapsed_time(e), nb / s.elapsed_time(e) / 1e3))
print("ok")
