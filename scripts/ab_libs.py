"""Same-box A/B of the library the process loads (WFOT_LIB_PATH): fused misfit + gradient and misfit-only rates on the
three window shapes, CUDA events, best of 4.  usage: WFOT_LIB_PATH=<so> python scripts/ab_libs.py"""
import os
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
import _inputs as I
from waveform_ot_b200 import _cabi as C
from waveform_ot_b200 import batch as B

SHAPES = {"cfg5": (1024, 256, 256, 0.04, 9472), "cfg1": (256, 80, 512, 0.03, 8192), "cfg4": (61, 79, 61, 0.04, 61440)}


def ev_time(fn, reps=4):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


for name, (nt, nug, ntg, lam, nb) in SHAPES.items():
    w = torch.from_numpy(I.random_walk_windows(513, nt, seed=5)).cuda()
    w = w[torch.arange(nb + 1) % w.shape[0]].contiguous()
    t = torch.linspace(0, 1, nt, device="cuda")
    grid = (0.0, 1.0, -1.3, 1.3, nug, ntg)
    tg = B.Target.from_waveform(t, w[0], grid, nug, ntg, lam)
    g = B.pack_grids(grid)
    ws = torch.empty(C.lib.wfot_misfit_grad_workspace_bytes(nb, nt, nug, ntg), dtype=torch.uint8, device="cuda")
    for label, kw in (("misfit+grad W2", dict(distfunc="W2")), ("misfit only W2", dict(distfunc="W2", want_grad=False)),
                      ("misfit only W12", dict(distfunc="W12", want_grad=False))):
        try:
            fn = lambda: B.misfit_grad_batch(t, w[1:], g, nug, ntg, lam, tg, workspace=ws, **kw)
            r = fn(); torch.cuda.synchronize()
            ms = ev_time(fn)
            print("%s %s B=%d %-16s %.3f ms  %.1f evals/s  sumW %.12e" % (
                os.environ.get("WFOT_LIB_PATH", "default"), name, nb, label, ms, nb / ms * 1e3, r["W"].sum().item()), flush=True)
        except Exception as ex:
            print(name, label, "not supported by this build:", type(ex).__name__)
