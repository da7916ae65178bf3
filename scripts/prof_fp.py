"""ncu driver: a few launches of the materialising fingerprint kernel and/or the fused kernel.
usage: prof_fp.py <windows> <fp|fused|both> [cfg5|cfg1|cfg4]"""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
import _inputs as O
from waveform_ot_b200 import batch as B

SHAPES = {"cfg5": (1024, 256, 256, 0.04), "cfg1": (256, 80, 512, 0.03), "cfg4": (61, 79, 61, 0.04)}
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 37
which = sys.argv[2] if len(sys.argv) > 2 else "both"
nt, nug, ntg, lam = SHAPES[sys.argv[3] if len(sys.argv) > 3 else "cfg5"]
w = torch.from_numpy(O.random_walk_windows(min(nb, 64) + 1, nt, seed=5)).cuda()
w = w[torch.arange(nb + 1) % w.shape[0]].contiguous()
t = torch.linspace(0, 1, nt, device="cuda")
grid = (0.0, 1.0, -1.3, 1.3, nug, ntg)
tg = B.Target.from_waveform(t, w[0], grid, nug, ntg, lam)
for _ in range(2):
    if which in ("both", "fp"):
        B.fingerprint_batch(t, w[1:], grid, nug, ntg, lam, deriv=True)
    if which in ("both", "fused"):
        B.misfit_grad_batch(t, w[1:], grid, nug, ntg, lam, tg)
torch.cuda.synchronize()
print("ok")
