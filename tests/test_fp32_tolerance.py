"""CPU: the FP32 rounding tolerance tau32 of the pruned scan (csrc/wfot_device.cuh) against a NumPy emulation of
the kernel's FP32 operation sequence.

The scan keeps, per pixel, every segment whose FP32 squared distance is within tau32 of the FP32 minimum and lets
FP64 (reference operation order) pick the nearest among them, so the nearest-segment index is exact provided
|D32(s) - D(s)| <= tau32 / 2 for every segment s (then D32(s*) <= D32_min + tau32 for the true nearest s*).
DESIGN.md section 3.3 derives the bound; this test measures it: the same elementary operations - inputs rounded
to FP32, two chained FMAs per projection, saturating subtract, FMA of the squares - emulated in NumPy (an FMA of
FP32 operands is evaluated in FP64, where the product is exact, and rounded once) on two million random
(pixel, segment) pairs in the kernel's scaled frame, including the nearly-touching pairs where the relative error
of D is largest."""
import numpy as np


def tau32(d2):
    d2 = np.asarray(d2, dtype=np.float64)
    return 2.5000006e-6 * np.sqrt(d2) * (1 + 2.0 ** -22) + 3.0e-7 * d2 + 1.0e-12


def f32(x):
    return np.asarray(x, dtype=np.float32)


def fma32(a, b, c):
    return f32(a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64))


def d32_kernel(px, py, ax, ay, bx, by):
    """prep_window's table entries and scan_block / tile_mask's distance, op for op (FP64 in, FP32 arithmetic)."""
    cx, cy = bx - ax, by - ay
    ln = np.sqrt(cx * cx + cy * cy)
    ex, ey = cx / ln, cy / ln
    mx, my = ax + 0.5 * cx, ay + 0.5 * cy
    fex, fey = f32(ex), f32(ey)
    nam, nbm = f32(-(mx * ex + my * ey)), f32(-(mx * ey - my * ex))
    h = f32(0.5 * ln)
    fpx, fpy = f32(px), f32(py)
    P = fma32(fpx, fex, nam)
    Q = fma32(fpx, fey, nbm)
    al = fma32(fpy, fey, P)
    pe = fma32(fpy, -fex, Q)
    tm = np.clip(f32(np.abs(al) - h), np.float32(0), np.float32(1))
    return fma32(tm, tm, f32(pe * pe)).astype(np.float64)


def d_exact(px, py, ax, ay, bx, by):
    cx, cy = bx - ax, by - ay
    lam = np.clip(((px - ax) * cx + (py - ay) * cy) / (cx * cx + cy * cy), 0.0, 1.0)
    dx, dy = px - ax - lam * cx, py - ay - lam * cy
    return dx * dx + dy * dy


def test_tau32_covers_fp32_rounding():
    rng = np.random.default_rng(2026)
    n = 2_000_000
    # scaled frame: everything inside a box of diagonal < 1 around the origin
    ax, ay = rng.uniform(-0.35, 0.35, n), rng.uniform(-0.35, 0.35, n)
    ang = rng.uniform(0, 2 * np.pi, n)
    ln = 10.0 ** rng.uniform(-4.5, -0.7, n)                     # segment lengths over four decades
    bx, by = ax + ln * np.cos(ang), ay + ln * np.sin(ang)
    px, py = rng.uniform(-0.35, 0.35, n), rng.uniform(-0.35, 0.35, n)
    # a third of the pixels sit next to their segment (distances down to 1e-7 of the frame)
    k = n // 3
    s = rng.uniform(-0.2, 1.2, k)
    off = 10.0 ** rng.uniform(-7, -2, k) * rng.choice([-1.0, 1.0], k)
    px[:k] = ax[:k] + s * (bx[:k] - ax[:k]) - off * np.sin(ang[:k])
    py[:k] = ay[:k] + s * (by[:k] - ay[:k]) + off * np.cos(ang[:k])
    D32 = d32_kernel(px, py, ax, ay, bx, by)
    D = d_exact(px, py, ax, ay, bx, by)
    err = np.abs(D32 - D)
    ratio = err / tau32(np.minimum(D32, D))
    assert ratio.max() <= 0.5, ("worst |D32 - D| / tau32 = %.3f" % ratio.max())
    # the margin actually used: report it so a change of the table layout or the op order shows up here
    assert ratio.max() >= 0.02            # the bound is not vacuous either (order-of-magnitude check)


def test_scan_skip_margin():
    """scan_block skips a tile when lb (1 - 2e-6) - 4e-6 > sqrt(b1): with lb a lower bound of the true distance of
    every segment of the tile, every FP32 distance of the tile then exceeds b1 + tau32(b1)."""
    d = 10.0 ** np.linspace(-7, 0, 2001)                         # sqrt(b1)
    lb = (d + 4.0e-6) * 1.0000023                                # smallest skipped lower bound (kernel's tq)
    worst32 = lb * lb - tau32(lb * lb) / 2                       # smallest FP32 value a segment at distance lb can take
    assert np.all(worst32 > d * d + tau32(d * d))
