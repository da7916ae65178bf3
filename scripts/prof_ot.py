"""Timing / ncu driver for the batched 1-D OT kernel (cfg2 shape), outputs preallocated, C ABI called directly.
usage: prof_ot.py <pairs> [reps] [bins]"""
import sys
import torch
sys.path.insert(0, ".")
from waveform_ot_b200 import _cabi as C
from waveform_ot_b200 import batch as B

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
dev = "cuda"
f = torch.rand(nb, n, device=dev) + 1e-3
g = torch.rand(nb, n, device=dev) + 1e-3
x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
W = torch.zeros(nb, 2, dtype=torch.float64, device=dev)
dpos = torch.zeros(nb, 2, dtype=torch.float64, device=dev)
dW1 = torch.empty(nb, n, dtype=torch.float64, device=dev)
dW2 = torch.empty(nb, n, dtype=torch.float64, device=dev)
amp = torch.empty(nb, dtype=torch.float64, device=dev)
st = B.Status()


def run(pmask, deriv):
    C.check(C.lib.wfot_ot1d_batch(C.ptr(f), C.ptr(g), C.F32, C.ptr(x), C.ptr(x), n, n, 0, 0, n, n, nb, pmask, deriv,
                                  C.ptr(W), C.ptr(dW1) if deriv else None, C.ptr(dW2) if deriv else None,
                                  C.ptr(dpos), C.ptr(amp), None, None, None, C.ptr(st.t), None))


s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, pm, dv, bytes_alg, bytes_moved in (("W12+dW1+dW2", 3, 1, 12 * n, 24 * n), ("W2+dW2", 2, 1, 12 * n, 16 * n),
                                             ("W12 only", 3, 0, 8 * n, 8 * n)):
    run(pm, dv); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        s.record(); run(pm, dv); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print("ot1d %-12s B=%d n=%d %.3f ms  %.2f Mpairs/s  %.1f Gknots/s  %.0f GB/s algorithmic (SURVEY 8d: 12 B/bin/pair)  %.0f GB/s moved" % (
        name, nb, n, best, nb / best / 1e3, nb * (2 * n - 1) / best / 1e6, nb * bytes_alg / best / 1e6, nb * bytes_moved / best / 1e6))
print("ok")
