import sys, numpy as np, torch
sys.path.insert(0, ".")
from waveform_ot_b200 import batch as B
n = 1024; nb = 1000000
g = torch.Generator(device="cuda").manual_seed(3)
f = torch.rand(nb, n, device="cuda", generator=g) + 0.05
t = torch.rand(nb, n, device="cuda", generator=g) + 0.05
# numpy.linspace (the reference's grids): the kernel recognises it and computes x instead of loading it;
# "torch" = torch.linspace, whose last bits differ: the general path
xs = {"numpy": torch.from_numpy(np.linspace(0, 1, n)).cuda(), "torch": torch.linspace(0, 1, n, device="cuda", dtype=torch.float64)}
which = sys.argv[1] if len(sys.argv) > 1 else "numpy"
x = xs[which]
for df in ("W2", "W12"):
    fn = lambda: B.ot1d_batch(f, t, x, x, distfunc=df, derivatives=True)
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize(); best = min(best, s.elapsed_time(e))
    print("%s grid, %s: %.3f ms %.2f M pairs/s" % (which, df, best, nb / best / 1e3), flush=True)
