import sys
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
import torch, _inputs as I
from waveform_ot_b200 import _cabi as C, batch as B
nt, nug, ntg, lam, nb = 1024, 256, 256, 0.04, 9472
w = torch.from_numpy(I.random_walk_windows(513, nt, seed=5)).cuda()
w = w[torch.arange(nb + 1) % w.shape[0]].contiguous()
t = torch.linspace(0, 1, nt, device="cuda")
grid = (0.0, 1.0, -1.3, 1.3, nug, ntg)
tg = B.Target.from_waveform(t, w[0], grid, nug, ntg, lam)
g = B.pack_grids(grid)
def ev(fn, reps=5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize(); best = min(best, s.elapsed_time(e))
    return best
for ov in (0, 1):
    for chunk in (0, 1895, 2368, 3158, 4736, 9472):
        C.lib.wfot_dev_set_option(C.OPT_PIPELINE, 2); C.lib.wfot_dev_set_option(C.OPT_OVERLAP, ov); C.lib.wfot_dev_set_option(C.OPT_SPLIT_CHUNK, chunk)
        ws = torch.empty(C.lib.wfot_misfit_grad_workspace_bytes(nb, nt, nug, ntg), dtype=torch.uint8, device="cuda")
        fn = lambda: B.misfit_grad_batch(t, w[1:], g, nug, ntg, lam, tg, workspace=ws)
        fn(); ms = ev(fn)
        print("overlap %s chunk %5d: %.3f ms %.0f evals/s  ws %.2f GB" % ("on " if ov == 0 else "off", chunk, ms, nb / ms * 1e3, ws.numel() / 1e9), flush=True)
