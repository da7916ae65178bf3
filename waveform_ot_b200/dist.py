"""Multi-GPU sharding of independent waveform windows.

Windows (stations x components x trial models) are independent units: each rank takes a
contiguous range, runs the fused kernel on it, reduces locally to [sum misfit, sum gradient]
with a fixed summation order (the misfits are bit-reproducible; the per-window gradient rows
come out of FP64 L2 reductions and carry ~1e-15 relative run-to-run noise), and ONE allreduce (NCCL over NVLink on GPUs; gloo in the CPU
tests) combines the ranks.  There is no other data-path collective.
"""
from __future__ import annotations

import os


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous split of n_items over world ranks; the first (n_items % world) ranks get one more."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """One process per GPU (torchrun).  backend defaults to nccl on CUDA, gloo otherwise."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world == 1 or dist.is_initialized():
        return rank, world, local
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group(backend, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    return rank, world, local


def allreduce_sum_(vec):
    """In-place sum over ranks of the packed [sum misfit, sum dwg, sum grad] vector."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec


def pack_local_sums(W, dwg, grad, reducer):
    """(B,2), (B,), (B,2,nt) -> (3 + 2*nt,) = [sum W^t, sum W^u, sum dwg, sum grad_t, sum grad_u];
    `reducer(x)` sums over the leading axis with a fixed order (batch.sum_windows on the GPU)."""
    import torch
    B = W.shape[0]
    packed = torch.cat([W.reshape(B, 2), dwg.reshape(B, 1), grad.reshape(B, -1)], dim=1).contiguous()
    return reducer(packed)


def sharded_misfit_grad(evaluate, n_windows, reducer):
    """Run `evaluate(lo, hi) -> (W, dwg, grad)` on this rank's contiguous shard of n_windows and
    return the allreduced packed sums (identical on every rank)."""
    rank, world, _ = env_rank_world()
    lo, hi = shard_bounds(n_windows, rank, world)
    W, dwg, grad = evaluate(lo, hi)
    return allreduce_sum_(pack_local_sums(W, dwg, grad, reducer))


def _shard_rows(x, n, lo, hi, what):
    """Rows of a per-window / shared / periodic array (leading axis) for the shard [lo, hi) of n windows.
    The kernels index such rows with the SHARD-LOCAL window number (b % rows), so a periodic layout (one row
    per station/component, repeated for every trial model) is rotated by lo % rows: local window 0 then finds
    the row of global window lo."""
    rows = int(x.shape[0])
    if rows == 1:
        return x
    if rows == n:
        return x[lo:hi].contiguous()
    if n % rows != 0:
        raise ValueError("%s: %d rows for %d windows (neither one per window nor a period of the batch)" % (what, rows, n))
    k = lo % rows
    if k == 0:
        return x
    import torch
    return torch.roll(x, shifts=-k, dims=0).contiguous()


def misfit_grad_sharded(t, w, grids, nug, ntg, lambdav, target, **kw):
    """The multi-GPU evaluation of a stacked misfit: this rank runs the fused kernel on its contiguous shard
    of the windows `w` (B, nt) (a per-window time axis `t` (B, nt) is sharded alike), reduces it to
    [sum W^t, sum W^u, sum dW^t/dx0, sum dW^t/dw, sum dW^u/dw] with the fixed-order window sum, and the ranks
    combine with ONE allreduce.  Returns the packed (3 + 2 nt,) device vector, identical on every rank.
    Grids / observed windows with one row per window are sharded alike; periodic rows (one per
    station/component, window b uses row b % rows) are rotated so that every shard starts on its own phase;
    a rank whose shard is empty (fewer windows than ranks) contributes zeros."""
    import torch
    from . import batch as B
    rank, world, _ = env_rank_world()
    n = w.shape[0]
    nt = w.shape[1]
    lo, hi = shard_bounds(n, rank, world)
    if hi == lo:
        dev = w.device if isinstance(w, torch.Tensor) else B._device()
        return allreduce_sum_(torch.zeros(3 + 2 * nt, dtype=torch.float64, device=dev))
    tt = t[lo:hi] if getattr(t, "ndim", 1) == 2 else t
    g = grids
    if hasattr(grids, "shape"):
        g = _shard_rows(grids, n, lo, hi, "grids")
    tg = target
    if target.rows > 1:
        tg = B.Target(*[_shard_rows(x, n, lo, hi, "target") for x in (target.cdf_t, target.x_t, target.cdf_u, target.x_u)])
    r = B.misfit_grad_batch(tt, w[lo:hi], g, nug, ntg, lambdav, tg, **kw)
    return allreduce_sum_(pack_local_sums(r["W"], r["dwg"], r["grad"], B.sum_windows))
