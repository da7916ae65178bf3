"""Single-evaluation latency breakdown (cfg1 shape): fused kernel alone (CUDA events) vs the whole adapter call."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from waveform_ot_b200 import batch as B, adapters, _cabi as C

def ricker_obs():
    f = 1.0 * 25 * 4 / 128
    t = np.arange(-2.0, (4 - 4 / 128) / 2, 4 / 128)
    w = (1.0 - 2.0 * np.pi ** 2 * f ** 2 * t ** 2) * np.exp(-np.pi ** 2 * f ** 2 * t ** 2)
    return np.linspace(-2.0, 2.0, 256), 1.6 * np.concatenate((w, w))

to, wo = ricker_obs()
grid = (-2.0, 2.0, -1.8, 4.2, 80, 512)
tg = adapters.make_target(to, wo, grid, 0.03)
fw = B.ricker_batch(np.array([[0.7, 1.3, 0.8]]), (-2.0, 2.0), deriv=True)
g = B.pack_grids(grid)
ws = torch.empty(C.lib.wfot_misfit_grad_workspace_bytes(1, 256, 80, 512), dtype=torch.uint8, device="cuda")
res = None
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for cl in ("8", "4", "2", "1"):
    os.environ["WFOT_DEV_CLUSTER"] = cl
    for _ in range(3):
        res = B.misfit_grad_batch(fw["t"], fw["w"], g, 80, 512, 0.03, tg, workspace=ws, out=res)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        s.record(); res = B.misfit_grad_batch(fw["t"], fw["w"], g, 80, 512, 0.03, tg, workspace=ws, out=res); e.record()
        torch.cuda.synchronize(); best = min(best, s.elapsed_time(e))
    data = [tg, "W2", (-2.0, 2.0), grid, 0.03, False, 0.5, 45.0]
    X1 = np.array([[0.7, 1.3, 0.8]])
    adapters.optfunc_ricker_batch(X1, data)
    t0 = time.perf_counter()
    for _ in range(50):
        adapters.optfunc_ricker_batch(X1, data)
    print("cluster<=%s: fused call (events) %.3f ms; whole adapter call %.3f ms" % (cl, best, (time.perf_counter() - t0) / 50 * 1e3))
