import sys, torch
sys.path.insert(0, ".")
exec(open("scripts/quick_perf.py").read().split("def main")[0])
import os
def run(name, nt, nug, ntg, nb, lam):
    w = torch.from_numpy(O.random_walk_windows(min(nb, 64) + 1, nt, seed=5)).cuda()
    w = w[torch.arange(nb + 1) % w.shape[0]].contiguous()
    t = torch.linspace(0, 1, nt, device="cuda")
    grid = (0.0, 1.0, -1.3, 1.3, nug, ntg)
    tg = B.Target.from_waveform(t, w[0], grid, nug, ntg, lam)
    g = B.pack_grids(grid)
    ws = torch.empty(C.lib.wfot_misfit_grad_workspace_bytes(nb, nt, nug, ntg), dtype=torch.uint8, device="cuda")
    fn = lambda: B.misfit_grad_batch(t, w[1:], g, nug, ntg, lam, tg, workspace=ws)
    fn(); torch.cuda.synchronize()
    ms = ev_time(fn)
    print("%s NT=%s: B=%d %.2f ms %.0f windows/s" % (name, os.environ.get("WFOT_DEV_NT", "auto"), nb, ms, nb / ms * 1e3))
for nt_ in ("64", "128", "256"):
    os.environ["WFOT_DEV_NT"] = nt_
    run("cfg4", 61, 79, 61, 30 * 1024, 0.04)
    run("cfg1", 256, 80, 512, 2368, 0.03)
    run("cfg5", 1024, 256, 256, 1184, 0.04)
