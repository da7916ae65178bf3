"""A/B on the GPU: single-kernel vs two-kernel (scan + resolve) form of the fused path.
Checks that both forms give the same results, then times them (CUDA events, best of 3).
usage: ab_split.py [cfg5|cfg1|cfg4] [windows]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
import _inputs as I
from waveform_ot_b200 import _cabi as C
from waveform_ot_b200 import batch as B

SHAPES = {"cfg5": (1024, 256, 256, 0.04, 9472), "cfg1": (256, 80, 512, 0.03, 8192), "cfg4": (61, 79, 61, 0.04, 61440)}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
nt, nug, ntg, lam, nb = SHAPES[name]
if len(sys.argv) > 2:
    nb = int(sys.argv[2])


def opt(i, v):
    C.lib.wfot_dev_set_option(i, v)


def ev_time(fn, reps=3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


w = torch.from_numpy(I.random_walk_windows(min(nb, 512) + 1, nt, seed=5)).cuda()
w = w[torch.arange(nb + 1) % w.shape[0]].contiguous()
t = torch.linspace(0, 1, nt, device="cuda")
grid = (0.0, 1.0, -1.3, 1.3, nug, ntg)
tg = B.Target.from_waveform(t, w[0], grid, nug, ntg, lam)
g = B.pack_grids(grid)


def run(keep=False):
    ws = torch.empty(C.lib.wfot_misfit_grad_workspace_bytes(nb, nt, nug, ntg), dtype=torch.uint8, device="cuda")
    fn = lambda: B.misfit_grad_batch(t, w[1:], g, nug, ntg, lam, tg, workspace=ws)
    r = fn()
    torch.cuda.synchronize()
    ms = ev_time(fn)
    st = r["status"].read()
    return ms, r, st


opt(C.OPT_PIPELINE, 1)
ms0, r0, st0 = run()
print("%s B=%d  single-kernel: %.3f ms  %.1f evals/s  status %s" % (name, nb, ms0, nb / ms0 * 1e3, st0[:5].tolist()), flush=True)
W0, G0, D0 = r0["W"].clone(), r0["grad"].clone(), r0["dwg"].clone()
# (label, overlap[0 two streams / 1 sequential], resolve shape [0 auto, 1: 128 regs x2/SM, 2: 80 regs x3/SM, 5: 128 threads x6/SM], chunk, scan shape [0: 80 regs x3, 2: 128 x2])
VARIANTS = [("two-kernel default", 0, 0, 0, 0),
            ("two-kernel, resolve 128 regs x2", 0, 1, 0, 0),
            ("two-kernel, resolve 80 regs x3", 0, 2, 0, 0),
            ("two-kernel, scan 128 regs x2", 0, 0, 0, 2),
            ("two-kernel, scan 256 threads x3", 0, 0, 0, 3),
            ("two-kernel, scan 128 threads x6", 0, 0, 0, 4),
            ("two-kernel, resolve 128 threads x6", 0, 5, 0, 0),
            ("two-kernel sequential one chunk", 1, 0, nb, 0),
            ("two-kernel chunk 32/SM", 0, 0, 4736, 0)]
if len(sys.argv) > 3:
    VARIANTS = [v for i, v in enumerate(VARIANTS) if str(i) in sys.argv[3].split(",")]
for label, ov, shape, chunk, sc in VARIANTS:
    opt(C.OPT_PIPELINE, 2); opt(C.OPT_OVERLAP, ov); opt(C.OPT_RESOLVE_SHAPE, shape); opt(C.OPT_SPLIT_CHUNK, chunk)
    opt(C.OPT_SCAN_SHAPE, sc)
    ms, r, st = run()
    dW = (r["W"] - W0).abs().max().item()
    dD = (r["dwg"] - D0).abs().max().item()
    dG = ((r["grad"] - G0).abs().max() / G0.abs().max()).item()
    print("%s B=%d  %s: %.3f ms  %.1f evals/s  max|dW| %.3g max|ddwg| %.3g max rel dgrad %.3g status %s" % (
        name, nb, label, ms, nb / ms * 1e3, dW, dD, dG, st[:5].tolist()), flush=True)
for i in range(8):
    opt(i, 0)
