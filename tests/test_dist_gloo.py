"""CPU, world_size = 2, gloo: the multi-GPU path's host logic (contiguous window shards, fixed-order
local reduction, ONE allreduce of [sum misfit, sum dwg, sum gradient]) gives the single-process
result.  The per-window evaluation is played by the oracle here (no GPU in this container)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _windows():
    from oracle import wfot_oracle as O
    nt, grid, lam = 40, (0.0, 1.0, -1.3, 1.3, 12, 16), 0.05
    w = O.random_walk_windows(8, nt, seed=3).astype(np.float64)
    t = np.linspace(0, 1, nt)
    _, tgt = O.build_ot_from_waveform(t, w[0], grid, lambdav=lam)
    return O, t, w[1:], grid, lam, tgt


def _evaluate_factory():
    O, t, w, grid, lam, tgt = _windows()

    def evaluate(lo, hi):
        Ws, dgs, grads = [], [], []
        for b in range(lo, hi):
            W, dr, dg, _, _ = O.misfit_grad_window(t, w[b], grid, tgt, lambdav=lam)
            Ws.append(W); dgs.append(dg[0]); grads.append(np.stack(dr))
        nt = w.shape[1]
        return (torch.tensor(np.array(Ws).reshape(-1, 2)), torch.tensor(np.array(dgs).reshape(-1)),
                torch.tensor(np.array(grads).reshape(-1, 2, nt)))
    return evaluate, w.shape[0]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from waveform_ot_b200 import dist as wd
    wd.init_process_group("gloo")
    evaluate, n = _evaluate_factory()
    out = wd.sharded_misfit_grad(evaluate, n, reducer=lambda x: x.sum(dim=0))
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    evaluate, n = _evaluate_factory()
    from waveform_ot_b200 import dist as wd
    W, dwg, grad = evaluate(0, n)
    ref = wd.pack_local_sums(W, dwg, grad, lambda x: x.sum(dim=0)).numpy()
    np.testing.assert_allclose(res[0], ref, rtol=1e-12, atol=1e-15)
    np.testing.assert_array_equal(res[0], res[1])          # every rank holds the same totals
