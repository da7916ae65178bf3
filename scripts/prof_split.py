"""ncu driver for the two-kernel form: warm-up call + one profiled call of k_scan / k_resolve.
usage: prof_split.py <windows> [cfg5|cfg1|cfg4] [resolve shape 1..3] [pipeline 1|2]"""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
import _inputs as I
from waveform_ot_b200 import _cabi as C
from waveform_ot_b200 import batch as B

SHAPES = {"cfg5": (1024, 256, 256, 0.04), "cfg1": (256, 80, 512, 0.03), "cfg4": (61, 79, 61, 0.04)}
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 592
nt, nug, ntg, lam = SHAPES[sys.argv[2] if len(sys.argv) > 2 else "cfg5"]
C.lib.wfot_dev_set_option(C.OPT_PIPELINE, int(sys.argv[4]) if len(sys.argv) > 4 else 2)
C.lib.wfot_dev_set_option(C.OPT_RESOLVE_SHAPE, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
C.lib.wfot_dev_set_option(C.OPT_OVERLAP, 1)                 # one kernel at a time under the profiler
C.lib.wfot_dev_set_option(C.OPT_SCAN_SHAPE, int(sys.argv[5]) if len(sys.argv) > 5 else 0)
w = torch.from_numpy(I.random_walk_windows(nb + 1, nt, seed=5)).cuda()
t = torch.linspace(0, 1, nt, device="cuda")
grid = (0.0, 1.0, -1.3, 1.3, nug, ntg)
tg = B.Target.from_waveform(t, w[0], grid, nug, ntg, lam)
for _ in range(2):
    B.misfit_grad_batch(t, w[1:], grid, nug, ntg, lam, tg)
torch.cuda.synchronize()
print("ok")
