"""Quick on-GPU timing of the probe and the fused kernel (development aid, not the bench)."""
import ctypes
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
import _inputs as O
from waveform_ot_b200 import _cabi as C
from waveform_ot_b200 import batch as B


def ev_time(fn, reps=3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


def main():
    print("device", torch.cuda.get_device_name(0), "SMs", C.lib.wfot_device_sm_count(), "cc", C.lib.wfot_device_cc())
    sink = torch.zeros(4, device="cuda")
    for packed in (0, 1):
        ops = ctypes.c_double()
        fn = lambda: C.check(C.lib.wfot_fp32_peak_probe(packed, 4000, C.ptr(sink), ctypes.byref(ops), None))
        fn()
        ms = ev_time(fn)
        print("fp32 probe packed=%d: %.2f ms  %.2f TFLOP/s (2 flop per fma)" % (packed, ms, 2 * ops.value / ms / 1e9))
    for name, nt, nug, ntg, nb, lam in (("cfg5", 1024, 256, 256, 592, 0.04), ("cfg1", 256, 80, 512, 1184, 0.03),
                                        ("cfg4", 61, 79, 61, 30 * 512, 0.04)):
        w = torch.from_numpy(O.random_walk_windows(min(nb, 64) + 1, nt, seed=5)).cuda()
        w = w[torch.arange(nb + 1) % w.shape[0]].contiguous()
        t = torch.linspace(0, 1, nt, device="cuda")
        grid = (0.0, 1.0, -1.3, 1.3, nug, ntg)
        tg = B.Target.from_waveform(t, w[0], grid, nug, ntg, lam)
        g = B.pack_grids(grid)
        ws = torch.empty(C.lib.wfot_misfit_grad_workspace_bytes(nb, nt, nug, ntg), dtype=torch.uint8, device="cuda")
        fn = lambda: B.misfit_grad_batch(t, w[1:], g, nug, ntg, lam, tg, workspace=ws)
        r = fn(); torch.cuda.synchronize()
        ms = ev_time(fn)
        pairs = nb * nug * ntg * (nt - 1)
        print("%s fused: B=%d %.2f ms  %.1f windows/s  %.3f Tpair/s  alg %.1f TFLOP/s  slow_px/window %.1f  executed pairs %.3f" % (
            name, nb, ms, nb / ms * 1e3, pairs / ms / 1e9, 15 * pairs / ms / 1e9,
            r["status"].read()[4] / nb, r["status"].scan_pairs() / pairs))
        d32 = torch.empty((nb, nug, ntg), dtype=torch.float32, device="cuda")
        fns = lambda: C.check(C.lib.wfot_scan_probe(C.ptr(t), C.ptr(w[1:]), 0, 0, nt, C.ptr(g), 1, nb, nug, ntg, C.ptr(d32), None))
        fns(); ms0 = ev_time(fns)
        print("   scan only: %.2f ms  %.3f Tpair/s  alg %.1f TFLOP/s" % (ms0, pairs / ms0 / 1e9, 15 * pairs / ms0 / 1e9))
        fnm = lambda: B.misfit_grad_batch(t, w[1:], g, nug, ntg, lam, tg, workspace=ws, want_grad=False)
        fnm(); ms2 = ev_time(fnm)
        print("   misfit only: %.2f ms" % ms2)
        nb2 = min(nb, 64)
        fn2 = lambda: B.fingerprint_batch(t, w[1:1 + nb2], g, nug, ntg, lam, deriv=True)
        fn2(); ms3 = ev_time(fn2)
        print("   materialising fingerprint B=%d: %.2f ms (%.1f windows/s)" % (nb2, ms3, nb2 / ms3 * 1e3))
    # cfg2: 1-D OT
    n, nb = 1024, 20000
    f = torch.rand(nb, n, device="cuda") + 1e-3
    gg = torch.rand(nb, n, device="cuda") + 1e-3
    x = torch.linspace(0, 1, n, dtype=torch.float64, device="cuda")
    fn = lambda: B.ot1d_batch(f, gg, x, x, "W12", derivatives=True)
    fn(); ms = ev_time(fn)
    print("cfg2 ot1d: B=%d %.2f ms  %.3f Mpairs/s  %.1f GB/s algorithmic" % (nb, ms, nb / ms / 1e3, nb * 12288 / ms / 1e6))


if __name__ == "__main__":
    main()
