"""cfg4 end to end (BASELINE.json configs[3] at its named size): 4096 trial models x 30 windows of 61 samples through
adapters.misfit_grad_models with pinned host inputs; wall-clock per call for both forms of the fused path."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from waveform_ot_b200 import _cabi as C
from waveform_ot_b200 import adapters

rng = np.random.default_rng(1)
M4, nr4, nc4, nt4 = 4096, 10, 3, 61
t4 = np.arange(float(nt4))
pulse = lambda sh, wd: np.exp(-0.5 * ((t4 - sh) / wd) ** 2) * np.sin(0.35 * (t4 - sh))
obs4 = np.stack([[pulse(22 + 2 * i + j, 4.0) for j in range(nc4)] for i in range(nr4)]) * 1e-3
obs4 += 2e-5 * rng.standard_normal(obs4.shape)
sh = rng.integers(-4, 5, size=M4)
base4 = np.stack([[pulse(22 + 2 * i + j, 4.0) for j in range(nc4)] for i in range(nr4)]) * 1e-3
pred4 = np.stack([np.roll(base4, int(s_), axis=-1) for s_ in sh]) * rng.uniform(0.7, 1.3, size=(M4, 1, 1, 1))
pred4_pin = torch.from_numpy(pred4).pin_memory()
J4_pin = torch.randn((M4, 9, nr4 * nc4 * nt4), dtype=torch.float64).pin_memory()
grids4 = adapters.buildFingerprintwindows(t4, obs4)
tg4 = adapters.make_targets_models(t4, obs4, grids4, 0.04)
for pipeline in (0,):
    C.lib.wfot_dev_set_option(C.OPT_PIPELINE, pipeline)
    for cm in (512, 768, 1024, 1366, 2048):
        adapters.misfit_grad_models(t4, pred4_pin[:64], grids4, tg4, 0.04, J=J4_pin[:64])
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            r = adapters.misfit_grad_models(t4, pred4_pin, grids4, tg4, 0.04, J=J4_pin, chunk_models=cm)
            best = min(best, time.perf_counter() - t0)
        print("pipeline %d chunk_models %4d: %.1f ms  %.0f models/s  (mis[0] %.6e)" % (pipeline, cm, best * 1e3, M4 / best, r[0][0]), flush=True)
C.lib.wfot_dev_set_option(C.OPT_PIPELINE, 0)
