import sys, torch
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
import _inputs as O
from waveform_ot_b200 import _cabi as C, batch as B
nt, nug, ntg = 1024, 256, 256
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 296
w = torch.from_numpy(O.random_walk_windows(64, nt, seed=5)).cuda()
w = w[torch.arange(nb) % 64].contiguous()
t = torch.linspace(0, 1, nt, device="cuda")
g = B.pack_grids((0.0, 1.0, -1.3, 1.3, nug, ntg))
d32 = torch.empty((nb, nug, ntg), dtype=torch.float32, device="cuda")
for _ in range(3):
    C.check(C.lib.wfot_scan_probe(C.ptr(t), C.ptr(w), 0, 0, nt, C.ptr(g), 1, nb, nug, ntg, C.ptr(d32), None))
torch.cuda.synchronize()
print("ok")
