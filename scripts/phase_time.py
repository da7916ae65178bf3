"""Time the two kernels of the two-kernel form separately (one chunk, sequential): full pass, scan only, resolve only
(on the scan results the full pass left in the workspace), resolve without the gradient phase.
usage: phase_time.py [cfg5|cfg1|cfg4] [windows] [resolve shape]"""
import sys

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
shape = int(sys.argv[3]) if len(sys.argv) > 3 else 0


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
ws = torch.empty(C.lib.wfot_misfit_grad_workspace_bytes(nb, nt, nug, ntg), dtype=torch.uint8, device="cuda")
opt(C.OPT_PIPELINE, 2); opt(C.OPT_OVERLAP, 1); opt(C.OPT_SPLIT_CHUNK, nb); opt(C.OPT_RESOLVE_SHAPE, shape)
for label, skip, grad in (("scan + resolve", 0, True), ("scan only", 1, True), ("resolve only", 2, True),
                          ("scan + resolve, no gradient", 0, False), ("resolve only, no gradient", 2, False)):
    opt(C.OPT_SKIP_KERNEL, 0)
    fn = lambda: B.misfit_grad_batch(t, w[1:], g, nug, ntg, lam, tg, workspace=ws, want_grad=grad)
    fn(); torch.cuda.synchronize()           # fresh scan results in the workspace
    opt(C.OPT_SKIP_KERNEL, skip)
    ms = ev_time(fn)
    print("%s B=%d shape %d  %-30s %8.3f ms  %.0f windows/s" % (name, nb, shape, label, ms, nb / ms * 1e3), flush=True)
for i in range(9):
    opt(i, 0)
