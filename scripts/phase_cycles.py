"""Per-phase SM cycles of k_resolve (wfot_dev_phase_cycles): window preparation, P1, P2 + P3, P4.
usage: phase_cycles.py [cfg5|cfg1|cfg4] [windows] [resolve shape]"""
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
w = torch.from_numpy(I.random_walk_windows(min(nb, 512) + 1, nt, seed=5)).cuda()
w = w[torch.arange(nb + 1) % w.shape[0]].contiguous()
t = torch.linspace(0, 1, nt, device="cuda")
grid = (0.0, 1.0, -1.3, 1.3, nug, ntg)
tg = B.Target.from_waveform(t, w[0], grid, nug, ntg, lam)
g = B.pack_grids(grid)
ws = torch.empty(C.lib.wfot_misfit_grad_workspace_bytes(nb, nt, nug, ntg), dtype=torch.uint8, device="cuda")
C.lib.wfot_dev_set_option(C.OPT_PIPELINE, 2); C.lib.wfot_dev_set_option(C.OPT_RESOLVE_SHAPE, shape)
fn = lambda: B.misfit_grad_batch(t, w[1:], g, nug, ntg, lam, tg, workspace=ws)
fn(); torch.cuda.synchronize()
cyc = torch.zeros(8, dtype=torch.int64, device="cuda")
C.lib.wfot_dev_phase_cycles(cyc.data_ptr())
fn(); torch.cuda.synchronize()
C.lib.wfot_dev_phase_cycles(None)
c = cyc.cpu().tolist()
tot = sum(c[:4])
print("%s B=%d shape %d: windows %d, cycles per window per CTA: prep %.0f (%.1f%%)  P1 %.0f (%.1f%%)  P2+P3 %.0f (%.1f%%)  P4 %.0f (%.1f%%)  total %.0f" % (
    name, nb, shape, c[4], c[0] / c[4], 100 * c[0] / tot, c[1] / c[4], 100 * c[1] / tot, c[2] / c[4], 100 * c[2] / tot,
    c[3] / c[4], 100 * c[3] / tot, tot / c[4]))
