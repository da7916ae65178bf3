"""Synthetic inputs for the development scripts (kept independent of oracle/: the oracle is test infrastructure)."""
import numpy as np


def random_walk_windows(B, nt, seed=5, dtype=np.float32):
    """cfg5 input rule: cumulative-sum random walk, moving average 8, mean removed,
    scaled to max|w| = 1."""
    rng = np.random.default_rng(seed)
    x = np.cumsum(rng.standard_normal((B, nt + 7)), axis=1)
    k = np.ones(8) / 8.0
    y = np.stack([np.convolve(r, k, mode="valid") for r in x])
    y -= y.mean(axis=1, keepdims=True)
    y /= np.max(np.abs(y), axis=1, keepdims=True)
    return y.astype(dtype)
