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

# ---- the same evaluation captured once in a CUDA graph and replayed (device-side sequence: forward model ->
#      fused misfit+gradient -> chain rule), inputs written into a static device buffer, results read back
os.environ["WFOT_DEV_CLUSTER"] = "8"
alpha = 0.5
Xd = torch.tensor([[0.7, 1.3, 0.8]], dtype=torch.float64, device="cuda")
out_host = torch.empty(4, dtype=torch.float64).pin_memory()
st = B.Status()

def device_eval():
    fw = B.ricker_batch(Xd, (-2.0, 2.0), deriv=True)
    r = B.misfit_grad_batch(fw["t"], fw["w"], g, 80, 512, 0.03, tg, workspace=ws, status=st)
    W, gr = r["W"], r["grad"]
    w2 = alpha * W[:, 0] + (1 - alpha) * W[:, 1]
    dr = (alpha * gr[:, 0] + (1 - alpha) * gr[:, 1]).contiguous()
    deriv = B.chain_batch(fw["dw"], dr)
    deriv[:, 0] = alpha * r["dwg"] / 4.0
    return torch.cat([w2, deriv[0]])

side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        device_eval()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    res_static = device_eval()
torch.cuda.synchronize()

def graph_eval(x):
    Xd.copy_(torch.from_numpy(x), non_blocking=True)
    graph.replay()
    out_host.copy_(res_static, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out_host.numpy()

x = np.array([[0.7, 1.3, 0.8]])
ref = graph_eval(x).copy()
t0 = time.perf_counter()
for _ in range(200):
    graph_eval(x)
dt = (time.perf_counter() - t0) / 200 * 1e3
w2, d = adapters.optfunc_ricker_batch(x, [tg, "W2", (-2.0, 2.0), grid, 0.03, False, 0.5, 45.0])
print("CUDA-graph replay of one evaluation: %.3f ms; matches the adapter: %s" % (
    dt, np.allclose(ref, np.concatenate([w2, d[0]]), rtol=1e-12, atol=0)))
