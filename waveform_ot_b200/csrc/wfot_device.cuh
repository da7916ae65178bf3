// wfot_device.cuh -- device-side building blocks shared by the materialising
// fingerprint kernel and the fused misfit+gradient kernel (sm_100a).
//
// Layout of the work:
//   * prep_window(): FP64 non-dimensionalisation in the reference's operation
//     order (libs/FingerprintLib.py:90-113) -> pn64; then an FP32 "rotated
//     frame" segment table in window-local, power-of-two-scaled coordinates.
//   * scan_block<R, T>(): the hot loop.  A thread owns 2 pixel columns x R rows and
//     walks the segment tiles its warp's pixel footprint cannot rule out (exact
//     pruning on per-tile bounding boxes); per (pixel, segment) pair it spends 5 FP32 lane
//     operations (4 of them issued as packed FFMA2/FMUL2 over a PAIR OF ROWS, the
//     per-segment quantities entering as scalar-broadcast operands so that every
//     FFMA2 reads at most 4 registers - three distinct register pairs would cost
//     3 clk instead of 2 on the register banks) and half an FMNMX3:
//         along = (p - mid).e      perp = (p - mid).n
//         D     = perp^2 + sat(|along| - h)^2          (h = half segment length)
//     which is the clamped point-to-segment distance of
//     libs/FingerprintLib.py:256-259 written in the segment's own frame.
//     Only the running minimum per tile of T segments is kept.
//   * resolve_pixel(): re-walks the winning tile, and evaluates every segment
//     whose FP32 distance is within the rounding tolerance of the minimum in
//     FP64 with exactly the reference's sequence of elementary operations, so
//     the selected index is np.argmin's (first minimum) on the FP64 field.
//     A pixel whose second-best *tile* is also within tolerance is resolved by
//     a full-segment cooperative rescan (resolve_pixel_warp).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/wfot.h"

namespace wfot {

constexpr int kTilePad = 16;        // the segment table is padded to a multiple of this
constexpr int kTileMin = 8;         // smallest argmin tile (sizes the per-tile bounding-box array)
// The argmin tile size T (segments per tile) is a template parameter of the scan / resolve functions:
// 16 for long waveforms, 8 for short ones (fewer FP32 re-evaluations per pixel, finer pruning).
constexpr float kBig = 3.0e38f;    // "no distance yet"
constexpr float kPadD = 5.0f;      // squared distance produced by padding segments (> any real one)

// ------------------------------------------------------------------ packed FP32
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// Fire-and-forget FP64 addition in L2.  atomicAdd() with an unused result compiles to ATOMG (the value travels back
// and holds a scoreboard) inside divergent code; the PTX reduction is REDG everywhere.
__device__ __forceinline__ void red_add(double* p, double v) {
    asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// The segment table, tile boxes and tile keys always live in shared memory; the pointers reach the
// scan / resolve functions through structs, where the compiler loses the address space and falls back
// to generic loads (LD.E instead of LDS: longer latency, long-scoreboard tracking).  Telling it restores LDS.
#define WFOT_ASSUME_SHARED(p) __builtin_assume(__isShared(p))

// ------------------------------------------------------------------ per-window header
struct WinHdr {
    double T0, Tstep, Tlast;   // pixel time axis:      np.linspace(tlimnfp[0], tlimnfp[1], ntg)
    double U0, Ustep, Ulast;   // pixel amplitude axis: np.linspace(ulimnfp[0], ulimnfp[1], nug)
    double ccx, ccy;           // origin of the window-local FP32 frame (normalised units)
    double sigma;              // power-of-two scale of the FP32 frame
    double du;                 // u1 - u0 (raw amplitude box)
    double u0raw, u1raw;       // raw amplitude box before an arctan transform
    int degenerate;            // number of zero-length segments (zeroed by the caller before prep_window())
};

// FP32 segment table in the rotated frame (shared memory), stored per PAIR of segments (2p, 2p + 1) so that
// the per-pixel re-evaluation of a tile runs as packed FFMA2 over two segments and the scan reads a pair with
// two 16-byte loads:
//   A[p]            = {ex0, ex1, ey0, ey1}        e = unit direction
//   A[Spad / 2 + p] = {-am0, -am1, -bm0, -bm1}    am = mid.e, bm = mid x e (scaled)
struct SegTable {
    const float4* A;   // [Spad] = [Spad / 2 direction pairs | Spad / 2 offset pairs]
    const float* H;    // [Spad] half length (scaled)
    const float4* bbox;   // per tile of `tile` segments: {xlo, xhi, ylo, yhi} of its vertices (scaled frame)
    int S;             // real segments
    int Spad;          // padded to a multiple of kTilePad
    int tile;          // segments per tile (8 or 16)
};

// Pixel footprint of one warp in the scaled frame (warp-uniform): the scan skips every segment
// tile whose bounding box is farther from the footprint than the largest running minimum of the
// warp's pixels (exact pruning: a skipped tile cannot hold a nearest segment of any of them).
struct Footprint {
    float x0, x1, y0, y1;
};

__device__ __forceinline__ double lin_axis(double a0, double step, double alast, int i, int n) {
    // numpy.linspace: y = arange(n)*step + start ; y[-1] = stop
    if (i == n - 1 && n > 1) return alast;
    return __dadd_rn(__dmul_rn((double)i, step), a0);
}

// ------------------------------------------------------------------ FP32 pair kernel (scalar form)
__device__ __forceinline__ float eval32(const SegTable& tb, int s, float px, float py) {
    WFOT_ASSUME_SHARED(tb.A); WFOT_ASSUME_SHARED(tb.H);
    const float* e = reinterpret_cast<const float*>(tb.A + (s >> 1)) + (s & 1);
    const float* m = reinterpret_cast<const float*>(tb.A + (tb.Spad >> 1) + (s >> 1)) + (s & 1);
    const float ex = e[0], ey = e[2], nam = m[0], nbm = m[2];
    const float h = tb.H[s];
    const float P = __fmaf_rn(px, ex, nam);      // px*ex - am
    const float Q = __fmaf_rn(px, ey, nbm);      // px*ey - bm
    const float al = __fmaf_rn(py, ey, P);       // along  = (p - mid).e
    const float pe = __fmaf_rn(py, -ex, Q);      // perp   = (p - mid) x e
    const float tm = __saturatef(__fadd_rn(fabsf(al), -h));
    return __fmaf_rn(tm, tm, __fmul_rn(pe, pe));
}

// rounding tolerance of the FP32 squared distance (scaled frame, |coords| <= 0.5, D < 1);
// derivation in DESIGN.md section 3.3
// one MUFU.SQRT (relative error <= 2^-23) instead of the 8-instruction IEEE sequence: every use below is an
// upper bound with margins that dwarf it
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float tau32(float d2) {
    return 2.5000006e-6f * sqrt_approx(d2) + 3.0e-7f * d2 + 1.0e-12f;
}

// ------------------------------------------------------------------ FP64 reference-order evaluation
// libs/FingerprintLib.py:256-259 for one (pixel, segment): the same elementary
// operations in the same order, no FMA contraction.
__device__ __forceinline__ void eval64(const double2* __restrict__ pn, int s, double px, double py,
                                       double& D, double& lam) {
    const double2 a = pn[s];
    const double2 b = pn[s + 1];
    const double cx = __dsub_rn(b.x, a.x);                     // delta_n (:112)
    const double cy = __dsub_rn(b.y, a.y);
    const double L = __dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy));   // lsq_n (:113)
    const double bx = __dsub_rn(px, a.x);                      // b = p - x0 (:256)
    const double by = __dsub_rn(py, a.y);
    const double num = __dadd_rn(__dmul_rn(bx, cx), __dmul_rn(by, cy));
    // (:257) lam = clip(num / L, 0, 1).  The quotient is only needed when the projection falls inside the segment:
    // num <= 0 clips to 0 and num >= L clips to 1 whatever the rounding of the division (L > 0), so pixels whose
    // nearest point is a vertex - most pixels away from the waveform - skip the FP64 division (warp-uniformly for
    // whole warps of such pixels).  A zero-length segment has num = 0 -> 0, as before.
    double l = (num <= 0.0) ? 0.0 : 1.0;
    if (num > 0.0 && num < L) l = fmin(fmax(__ddiv_rn(num, L), 0.0), 1.0);   // np.clip
    const double dx = __dsub_rn(bx, __dmul_rn(cx, l));         // ds = b - c*lam (:258)
    const double dy = __dsub_rn(by, __dmul_rn(cy, l));
    D = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));       // (:259)
    lam = l;
}

struct PixelHit {
    double D;     // FP64 squared distance of the selected segment
    double lam;   // clipped segment parameter
    int s;        // selected segment (np.argmin first-minimum)
};

// Resolve one pixel from the scan's tile statistics.
//   b1/t1 = smallest tile minimum and its tile, b2/b3 = 2nd / 3rd smallest tile minimum.
// Every segment whose true FP64 distance could be minimal has an FP32 distance <= thr
// (thr = b1 + rounding tolerance), and therefore lives in a tile whose minimum is <= thr.
//   b2 > thr            : only tile t1 -> 16 FP32 re-evaluations;
//   b2 <= thr < b3      : exactly one other tile; if it is t1-1 or t1+1 (a vertex shared
//                         across the tile boundary - by far the common case) the three tiles
//                         are walked; otherwise, or if b3 <= thr, the caller queues the pixel
//                         for the all-segment rescan.
// The FP32 walk only builds a candidate bit mask; the FP64 reference-order evaluations run
// afterwards in ascending segment order (strict '<' keeps np.argmin's first-minimum rule),
// so lanes of a warp do not serialise on each other's candidates.
// Returns false if the pixel must go to the full rescan.
// FP32 re-evaluation of the T segments of one tile for one pixel: bit j set <=> D32(segment) <= thr.
// Padding segments evaluate to kPadD > thr, so no bounds check is needed.
template <int T>
__device__ __forceinline__ void tile_dists(const SegTable& tb, int tile, float px, float py, float (&d)[T]) {
    WFOT_ASSUME_SHARED(tb.A); WFOT_ASSUME_SHARED(tb.H);
    const float4* __restrict__ E = tb.A + tile * (T / 2);
    const float4* __restrict__ M = E + (tb.Spad >> 1);
    const float2* __restrict__ H2 = reinterpret_cast<const float2*>(tb.H + tile * T);
    const uint64_t px2 = pack2(px, px), py2 = pack2(py, py);
#pragma unroll
    for (int p = 0; p < T / 2; ++p) {            // two segments per step, the same FP32 operations as the scan
        const float4 e = E[p], m = M[p];
        const float2 h = H2[p];
        const uint64_t ex2 = pack2(e.x, e.y), ey2 = pack2(e.z, e.w), nex2 = pack2(-e.x, -e.y);
        const uint64_t P = ffma2(px2, ex2, pack2(m.x, m.y));
        const uint64_t Q = ffma2(px2, ey2, pack2(m.z, m.w));
        float al0, al1;
        unpack2(ffma2(py2, ey2, P), al0, al1);
        const uint64_t pe = ffma2(py2, nex2, Q);
        const uint64_t tm = pack2(__saturatef(__fadd_rn(fabsf(al0), -h.x)), __saturatef(__fadd_rn(fabsf(al1), -h.y)));
        unpack2(ffma2(tm, tm, fmul2(pe, pe)), d[2 * p], d[2 * p + 1]);
    }
}

template <int T>
__device__ __forceinline__ unsigned dists_mask(const float (&d)[T], float thr) {
    unsigned mask = 0u;
#pragma unroll
    for (int j = 0; j < T; ++j) mask |= (d[j] <= thr) ? (1u << j) : 0u;
    return mask;
}

template <int T>
__device__ __forceinline__ unsigned tile_mask(const SegTable& tb, int tile, float px, float py, float thr) {
    float d[T];
    tile_dists<T>(tb, tile, px, py, d);
    return dists_mask<T>(d, thr);
}

// FP64 reference-order evaluation of the candidate segments of one tile, ascending (strict '<'
// keeps np.argmin's first-minimum rule across calls made in ascending tile order).
// (Measured and dropped: evaluating candidates in consecutive pairs with a shared vertex load - two independent
// FP64 chains per pass - costs registers and an unused evaluation for lone candidates: -4 % on cfg5.)
template <int T>
__device__ __forceinline__ void eval_candidates(const double2* __restrict__ pn, int tile, unsigned mask,
                                                double px, double py, PixelHit& hit) {
    while (mask) {
        const int j = __ffs((int)mask) - 1;
        mask &= mask - 1u;
        double D, l;
        eval64(pn, tile * T + j, px, py, D, l);
        if (D < hit.D) { hit.D = D; hit.lam = l; hit.s = tile * T + j; }
    }
}

// `other` = another tile holds a segment within the tolerance (b2 <= thr), `many` = two or more do (b3 <= thr).
template <int T>
__device__ __forceinline__ bool resolve_pixel_flagged(const SegTable& tb, const double2* __restrict__ pn,
                                                      float pxl, float pyl, double px, double py,
                                                      float thr, int t1, bool other, bool many, PixelHit& hit) {
    hit.D = CUDART_INF; hit.lam = 0.0; hit.s = t1 * T;
    if (!other) {                                     // the common case: one tile
        eval_candidates<T>(pn, t1, tile_mask<T>(tb, t1, pxl, pyl, thr), px, py, hit);
        return true;
    }
    if (many) return false;
    // exactly one other tile holds a candidate: resolved here only if it is a neighbour of t1
    const int ntiles = tb.Spad / T;
    const unsigned m0 = (t1 > 0) ? tile_mask<T>(tb, t1 - 1, pxl, pyl, thr) : 0u;
    const unsigned m2 = (t1 + 1 < ntiles) ? tile_mask<T>(tb, t1 + 1, pxl, pyl, thr) : 0u;
    if ((m0 | m2) == 0u) return false;
    const unsigned m1 = tile_mask<T>(tb, t1, pxl, pyl, thr);
    if (m0) eval_candidates<T>(pn, t1 - 1, m0, px, py, hit);
    eval_candidates<T>(pn, t1, m1, px, py, hit);
    if (m2) eval_candidates<T>(pn, t1 + 1, m2, px, py, hit);
    return true;
}

// The same from the scan's 4-byte result (tile | flags) alone: the FP32 minimum b1 is the smallest of the winning
// tile's distances, which the re-evaluation produces anyway (same operations as the scan: the same bits).
template <int T>
__device__ __forceinline__ bool resolve_pixel_coded(const SegTable& tb, const double2* __restrict__ pn,
                                                    float pxl, float pyl, double px, double py,
                                                    int t1, bool other, bool many, PixelHit& hit, float& b1) {
    float d[T];
    tile_dists<T>(tb, t1, pxl, pyl, d);
    float mn = d[0];
#pragma unroll
    for (int j = 1; j < T; ++j) mn = fminf(mn, d[j]);
    b1 = mn;
    const float thr = mn + tau32(mn);
    const unsigned m1 = dists_mask<T>(d, thr);
    hit.D = CUDART_INF; hit.lam = 0.0; hit.s = t1 * T;
    if (!other) {                                     // the common case: one tile
        eval_candidates<T>(pn, t1, m1, px, py, hit);
        return true;
    }
    if (many) return false;
    const int ntiles = tb.Spad / T;
    const unsigned m0 = (t1 > 0) ? tile_mask<T>(tb, t1 - 1, pxl, pyl, thr) : 0u;
    const unsigned m2 = (t1 + 1 < ntiles) ? tile_mask<T>(tb, t1 + 1, pxl, pyl, thr) : 0u;
    if ((m0 | m2) == 0u) return false;
    if (m0) eval_candidates<T>(pn, t1 - 1, m0, px, py, hit);
    eval_candidates<T>(pn, t1, m1, px, py, hit);
    if (m2) eval_candidates<T>(pn, t1 + 1, m2, px, py, hit);
    return true;
}

template <int T>
__device__ __forceinline__ bool resolve_pixel(const SegTable& tb, const double2* __restrict__ pn,
                                              float pxl, float pyl, double px, double py,
                                              float b1, int t1, float b2, float b3, PixelHit& hit) {
    const float thr = b1 + tau32(b1);
    return resolve_pixel_flagged<T>(tb, pn, pxl, pyl, px, py, thr, t1, !(b2 > thr), b3 <= thr, hit);
}

// Full rescan of every segment by one thread (queue-overflow fallback).
static __device__ __noinline__ void resolve_pixel_full(const SegTable& tb, const double2* __restrict__ pn,
                                                float pxl, float pyl, double px, double py,
                                                float b1, PixelHit& hit) {
    const float thr = b1 + tau32(b1);
    hit.D = CUDART_INF; hit.lam = 0.0; hit.s = 0;
    for (int s = 0; s < tb.S; ++s) {
        const float d32 = eval32(tb, s, pxl, pyl);
        if (d32 <= thr) {
            double D, l;
            eval64(pn, s, px, py, D, l);
            if (D < hit.D) { hit.D = D; hit.lam = l; hit.s = s; }
        }
    }
}

// Full rescan by a whole warp (lanes stride over segments); result valid in every lane.
__device__ __forceinline__ void resolve_pixel_warp(const SegTable& tb, const double2* __restrict__ pn,
                                                   float pxl, float pyl, double px, double py,
                                                   float b1, PixelHit& hit) {
    const int lane = threadIdx.x & 31;
    const float thr = b1 + tau32(b1);
    double bD = CUDART_INF, bl = 0.0;
    int bs = 0x7fffffff;
    for (int s = lane; s < tb.S; s += 32) {
        const float d32 = eval32(tb, s, pxl, pyl);
        if (d32 <= thr) {
            double D, l;
            eval64(pn, s, px, py, D, l);
            if (D < bD) { bD = D; bl = l; bs = s; }   // ascending s per lane: first minimum
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double oD = __shfl_xor_sync(0xffffffffu, bD, off);
        const double ol = __shfl_xor_sync(0xffffffffu, bl, off);
        const int os = __shfl_xor_sync(0xffffffffu, bs, off);
        if (oD < bD || (oD == bD && os < bs)) { bD = oD; bl = ol; bs = os; }
    }
    hit.D = bD; hit.lam = bl; hit.s = (bs == 0x7fffffff) ? 0 : bs;
}

// ------------------------------------------------------------------ warp footprints
// The pixel grid is cut into warp footprints of FC x FR thread blocks (FC * FR = 32; a thread block
// is 2 columns x R rows).  FC is chosen per window so that the footprint is as square as possible
// in normalised units, which is what makes the pruning of scan_block() effective.
struct FootMap {
    int fcl;     // log2(FC)
    int nfc;     // footprints along the time axis
    int nfoot;   // footprints per window
};

// wx, wy: extents of the pixel time / amplitude axes in the scaled frame.
template <int R>
__host__ __device__ inline FootMap make_footmap(int ntg, int nug, float wx, float wy) {
    const int ncp = (ntg + 1) >> 1, nrg = (nug + R - 1) / R;
    const float dx = ntg > 1 ? wx / (float)(ntg - 1) : 0.f, dy = nug > 1 ? wy / (float)(nug - 1) : 0.f;
    int fcl = 4;
    if (dx > 0.f && dy > 0.f) {
        // 2 FC dx = R (32 / FC) dy  ->  FC = sqrt(16 R dy / dx), rounded in log2
        float v = 16.f * (float)R * dy / dx;
        fcl = 0;
        while (fcl < 5 && v >= 2.f) { v *= 0.25f; ++fcl; }    // fcl = round(0.5 log2 v)
    }
    while (fcl > 0 && (1 << (fcl - 1)) >= ncp) --fcl;          // never wider than the grid ...
    while (fcl < 5 && (32 >> (fcl + 1)) >= nrg) ++fcl;         // ... nor taller
    FootMap m;
    m.fcl = fcl;
    const int FC = 1 << fcl, FR = 32 >> fcl;
    m.nfc = (ncp + FC - 1) / FC;
    m.nfoot = m.nfc * ((nrg + FR - 1) / FR);
    return m;
}

// upper bound of FootMap::nfoot over every footprint shape (host-side grid sizing)
template <int R>
inline int max_footprints(int ntg, int nug) {
    const int ncp = (ntg + 1) >> 1, nrg = (nug + R - 1) / R;
    int best = 0;
    for (int fcl = 0; fcl <= 5; ++fcl) {
        const int FC = 1 << fcl, FR = 32 >> fcl;
        const int n = ((ncp + FC - 1) / FC) * ((nrg + FR - 1) / FR);
        best = n > best ? n : best;
    }
    return best;
}

// Lane's pixel block inside footprint f, and the footprint's bounding box.
struct LaneBlock {
    int cp, rg;       // column pair / row group (clamped into the grid: duplicates do valid, unused work)
    bool owns;        // false for a clamped duplicate
    Footprint fp;
};

template <int R>
__device__ __forceinline__ LaneBlock lane_block(const FootMap& m, int f, int lane, int ntg, int nug,
                                                const float* pxs, const float* pys) {
    const int ncp = (ntg + 1) >> 1, nrg = (nug + R - 1) / R;
    const int FC = 1 << m.fcl, FR = 32 >> m.fcl;
    const int fc = f % m.nfc, fr = f / m.nfc;
    const int cpi = fc * FC + (lane & (FC - 1)), rgi = fr * FR + (lane >> m.fcl);
    LaneBlock b;
    b.owns = (cpi < ncp) && (rgi < nrg);
    b.cp = min(cpi, ncp - 1);
    b.rg = min(rgi, nrg - 1);
    const int c0 = 2 * fc * FC, c1 = min(2 * min(fc * FC + FC - 1, ncp - 1) + 1, ntg - 1);
    const int r0 = fr * FR * R, r1 = min((fr * FR + FR) * R - 1, nug - 1);
    const float xa = pxs[c0], xb = pxs[c1], ya = pys[r0], yb = pys[r1];
    b.fp.x0 = fminf(xa, xb); b.fp.x1 = fmaxf(xa, xb);
    b.fp.y0 = fminf(ya, yb); b.fp.y1 = fmaxf(ya, yb);
    return b;
}

// ------------------------------------------------------------------ the hot loop
// Thread-owned pixel block: columns (px0, px1) x rows py[0..R).  Slot k = 2*r + c.
// On return b1[k] = min_s D, t1[k] = tile holding it, b2[k] / b3[k] = 2nd / 3rd smallest tile minimum
// among the tiles that were evaluated.
//
// Exact pruning (must be called by all 32 lanes of a converged warp; `fp` is the bounding box of
// the warp's pixels).  Tiles are visited BEST FIRST: in the order of the distance between their
// bounding box and the footprint box, so the running minima tighten as early as possible.  Every lane
// keeps the keys (distance bits | tile index) of the tiles it owns (tile % 32 == lane) in `keys`
// (shared memory, >= Spad / T entries per warp); one warp minimum (REDUX) pops the next tile.
// Every lane tests a popped tile against ITS OWN pixel block: the tile is of no interest to the lane
// when the distance `lb` between the tile's bounding box and the block's satisfies
// lb (1 - 2e-6) - 4e-6 > sqrt(max over the lane's pixels of the running minimum b1): every FP32
// distance of such a tile exceeds b1 + tau32(b1) of every pixel of the lane (the FP32 rounding
// tolerance is 1.25e-6 absolute + 1.5e-7 relative in distance units, see tau32), so the tile can
// neither hold the FP64 nearest segment nor a near-tie the resolve step has to look at.  The warp
// skips a tile that no lane is interested in (one vote), and stops as soon as the popped key - a
// lower bound of the box distance of every remaining tile to every lane's block, which lies inside
// the footprint - fails that test for all lanes.
template <int R, int T>
__device__ __forceinline__ void scan_block(const SegTable& tb, const Footprint& fp, float px0, float px1,
                                           const float (&py)[R],
                                           float (&b1)[2 * R], int (&t1)[2 * R], float (&b2)[2 * R],
                                           float (&b3)[2 * R], int& tiles_done, unsigned* __restrict__ keys) {
    static_assert(R % 2 == 0, "rows are processed in pairs");
    WFOT_ASSUME_SHARED(tb.A); WFOT_ASSUME_SHARED(tb.H); WFOT_ASSUME_SHARED(tb.bbox); WFOT_ASSUME_SHARED(keys);
    const uint64_t px2 = pack2(px0, px1);
    uint64_t y2[R / 2];
#pragma unroll
    for (int i = 0; i < R / 2; ++i) y2[i] = pack2(py[2 * i], py[2 * i + 1]);
#pragma unroll
    for (int k = 0; k < 2 * R; ++k) { b1[k] = kBig; b2[k] = kBig; b3[k] = kBig; t1[k] = 0; }
    const int ntiles = tb.Spad / T;
    const int lane = threadIdx.x & 31;
    // keys: squared box distance to the footprint (low bits dropped: still a lower bound) | tile index
    const unsigned idxmask = ntiles > 1 ? (0xffffffffu >> __clz(ntiles - 1)) : 0u;
    // a lane owns the tiles lane, lane + 32, ...: the first kRegKeys of them live in registers, the rest in `keys`
    constexpr int kRegKeys = 4;
    constexpr unsigned kNoKey = 0xffffffffu;
    unsigned kr[kRegKeys];
    unsigned kmem = kNoKey;                      // minimum of the lane's keys kept in shared memory
#pragma unroll
    for (int i = 0; i < kRegKeys; ++i) kr[i] = kNoKey;
    for (int i = 0, tile = lane; tile < ntiles; ++i, tile += 32) {
        const float4 bb = tb.bbox[tile];
        const float dx = fmaxf(0.f, fmaxf(bb.x - fp.x1, fp.x0 - bb.y));
        const float dy = fmaxf(0.f, fmaxf(bb.z - fp.y1, fp.y0 - bb.w));
        const unsigned key = (__float_as_uint(__fmaf_rn(dy, dy, dx * dx)) & ~idxmask) | (unsigned)tile;
        if (i < kRegKeys) {
#pragma unroll
            for (int q = 0; q < kRegKeys; ++q) kr[q] = (q == i) ? key : kr[q];
        } else {
            keys[tile] = key;
            kmem = min(kmem, key);
        }
    }
    unsigned mykey = min(min(min(kr[0], kr[1]), min(kr[2], kr[3])), kmem);
    // the lane's pixel block in the scaled frame
    const float qx0 = fminf(px0, px1), qx1 = fmaxf(px0, px1);
    const float qy0 = fminf(py[0], py[R - 1]), qy1 = fmaxf(py[0], py[R - 1]);
    // thr2 = ((sqrt(lmax) + 4e-6) / (1 - 2e-6))^2 (+ 3e-7 relative for the approximate square root): a tile whose squared box distance exceeds it cannot matter
    float thr2 = kBig;   // from lmax = max over the lane's pixels of b1, updated per evaluated tile
    while (true) {
        const unsigned kmin = __reduce_min_sync(0xffffffffu, mykey);
        if (kmin == kNoKey) break;                                       // every tile visited
        if (__all_sync(0xffffffffu, __uint_as_float(kmin & ~idxmask) > thr2)) break;
        const int tile = (int)(kmin & idxmask);
        // the owner drops the key (branch-free for the register keys) and every lane re-derives its minimum
        if (ntiles <= 32) {                                              // short windows: one key per lane
            mykey = (mykey == kmin) ? kNoKey : mykey;
        } else {
#pragma unroll
            for (int q = 0; q < kRegKeys; ++q) kr[q] = (kr[q] == kmin) ? kNoKey : kr[q];
            if (kmem == kmin) {                                          // long windows only: rescan the lane's keys in memory
                keys[tile] = kNoKey;
                kmem = kNoKey;
                for (int t2 = lane + 32 * kRegKeys; t2 < ntiles; t2 += 32) kmem = min(kmem, keys[t2]);
            }
            mykey = min(min(min(kr[0], kr[1]), min(kr[2], kr[3])), kmem);
        }
        {
            const float4 bb = tb.bbox[tile];
            const float dx = fmaxf(0.f, fmaxf(bb.x - qx1, qx0 - bb.y));
            const float dy = fmaxf(0.f, fmaxf(bb.z - qy1, qy0 - bb.w));
            if (!__any_sync(0xffffffffu, !(__fmaf_rn(dy, dy, dx * dx) > thr2))) continue;
        }
        ++tiles_done;
        float tm[2 * R];
#pragma unroll
        for (int k = 0; k < 2 * R; ++k) tm[k] = kBig;
        const float4* __restrict__ E = tb.A + tile * (T / 2);
        const float4* __restrict__ M = E + (tb.Spad >> 1);
        const float* __restrict__ H = tb.H + tile * T;
#pragma unroll 2
        for (int j = 0; j < T; j += 2) {
            const float4 e = E[j >> 1], m = M[j >> 1];
            const float4 a0 = make_float4(e.x, e.z, m.x, m.z), a1 = make_float4(e.y, e.w, m.y, m.w);
            const float2 hh = *reinterpret_cast<const float2*>(H + j);
            // per column: P = px*ex - am, Q = px*ey - bm (both columns in one packed op)
            float P0[2], Q0[2], P1[2], Q1[2];
            unpack2(ffma2(px2, pack2(a0.x, a0.x), pack2(a0.z, a0.z)), P0[0], P0[1]);
            unpack2(ffma2(px2, pack2(a0.y, a0.y), pack2(a0.w, a0.w)), Q0[0], Q0[1]);
            unpack2(ffma2(px2, pack2(a1.x, a1.x), pack2(a1.z, a1.z)), P1[0], P1[1]);
            unpack2(ffma2(px2, pack2(a1.y, a1.y), pack2(a1.w, a1.w)), Q1[0], Q1[1]);
            const uint64_t ey0 = pack2(a0.y, a0.y), nex0 = pack2(-a0.x, -a0.x);
            const uint64_t ey1 = pack2(a1.y, a1.y), nex1 = pack2(-a1.x, -a1.x);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint64_t Pc0 = pack2(P0[c], P0[c]), Qc0 = pack2(Q0[c], Q0[c]);
                const uint64_t Pc1 = pack2(P1[c], P1[c]), Qc1 = pack2(Q1[c], Q1[c]);
#pragma unroll
                for (int rp = 0; rp < R / 2; ++rp) {
                    float al0, al1, bl0, bl1;
                    unpack2(ffma2(y2[rp], ey0, Pc0), al0, al1);        // along, segment j,   rows 2rp, 2rp+1
                    unpack2(ffma2(y2[rp], ey1, Pc1), bl0, bl1);        // along, segment j+1
                    const uint64_t pe0 = ffma2(y2[rp], nex0, Qc0);     // perp
                    const uint64_t pe1 = ffma2(y2[rp], nex1, Qc1);
                    const float u0 = __saturatef(__fadd_rn(fabsf(al0), -hh.x));
                    const float u1 = __saturatef(__fadd_rn(fabsf(al1), -hh.x));
                    const float v0 = __saturatef(__fadd_rn(fabsf(bl0), -hh.y));
                    const float v1 = __saturatef(__fadd_rn(fabsf(bl1), -hh.y));
                    const uint64_t u2 = pack2(u0, u1), v2 = pack2(v0, v1);
                    float d00, d01, d10, d11;
                    unpack2(ffma2(u2, u2, fmul2(pe0, pe0)), d00, d01);
                    unpack2(ffma2(v2, v2, fmul2(pe1, pe1)), d10, d11);
                    const int k0 = (2 * rp) * 2 + c, k1 = (2 * rp + 1) * 2 + c;
                    tm[k0] = fminf(tm[k0], fminf(d00, d10));
                    tm[k1] = fminf(tm[k1], fminf(d01, d11));
                }
            }
        }
        float mx = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * R; ++k) {
            const bool better = tm[k] < b1[k];
            b3[k] = fminf(b3[k], fmaxf(tm[k], b2[k]));
            b2[k] = fminf(b2[k], fmaxf(tm[k], b1[k]));
            t1[k] = better ? tile : t1[k];
            b1[k] = fminf(b1[k], tm[k]);
            mx = fmaxf(mx, b1[k]);
        }
        const float tq = (sqrt_approx(mx) + 4.0e-6f) * 1.0000023f;
        thr2 = tq * tq;
    }
}

// ------------------------------------------------------------------ window preparation
// Input sample loader (FP32 or FP64 waveform / time arrays).
__device__ __forceinline__ double load_sample(const void* p, int dtype, long long i) {
    return dtype == WFOT_F64 ? reinterpret_cast<const double*>(p)[i]
                             : (double)reinterpret_cast<const float*>(p)[i];
}

// Destination buffers may live in shared or global memory.
struct PrepOut {
    double2* pn;    // [nt]
    float4* A;      // [Spad]  pair layout, see SegTable
    float* H;       // [Spad]
    float4* bbox;   // [Spad / tile]  per-tile vertex bounding boxes (scaled frame); nullptr: not wanted
    int tile;       // segments per tile (8 or 16)
    float* pxs;     // [ntg]  scaled local pixel time coordinates
    float* pys;     // [nug]
    WinHdr* hdr;    // [1]
};

// One block prepares one window.  `red` = 64 doubles of shared scratch.
// transform != 0: arctan amplitude transform (libs/ricker_util.py:270-275) with the
// grid's (u0,u1); the amplitude box becomes (0,1) (libs/ricker_util.py:241-244).
__device__ __forceinline__ void prep_window(const void* t, const void* w, int dtype, long long toff,
                                            long long woff, int nt, const wfot_grid& g, int nug,
                                            int ntg, int transform, const PrepOut& o, double* red,
                                            double* pn_out /* nullable global (nt,2) */) {
    const int tid = threadIdx.x, nth = blockDim.x;
    double u0 = g.u0, u1 = g.u1;
    const double u0raw = g.u0, u1raw = g.u1;
    if (transform) { u0 = 0.0; u1 = 1.0; }
    const double delt = __dmul_rn(g.tantheta, __dsub_rn(g.t1, g.t0));      // :90
    const double du = __dsub_rn(u1, u0);
    double mnx = CUDART_INF, mxx = -CUDART_INF, mny = CUDART_INF, mxy = -CUDART_INF;
    for (int j = tid; j < nt; j += nth) {
        const double tj = load_sample(t, dtype, toff + j);
        double wj = load_sample(w, dtype, woff + j);
        if (transform) {   // un = 0.5 + arctan(((u-u0)+(u-u1))/(u1-u0))/pi
            const double up = __ddiv_rn(__dadd_rn(__dsub_rn(wj, u0raw), __dsub_rn(wj, u1raw)),
                                        __dsub_rn(u1raw, u0raw));
            wj = __dadd_rn(0.5, __ddiv_rn(atan(up), CUDART_PI));
        }
        double2 p;
        p.x = __ddiv_rn(__dsub_rn(tj, g.t0), delt);                        // :110
        p.y = __ddiv_rn(__dsub_rn(wj, u0), du);
        o.pn[j] = p;
        if (pn_out) { pn_out[2 * j] = p.x; pn_out[2 * j + 1] = p.y; }
        mnx = fmin(mnx, p.x); mxx = fmax(mxx, p.x);
        mny = fmin(mny, p.y); mxy = fmax(mxy, p.y);
    }
    {   // the four extrema through ONE pair of block barriers (exact whatever the order); o.pn visible below
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, off));
            mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, off));
            mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, off));
            mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, off));
        }
        __syncthreads();
        if (lane == 0) { red[wid] = mnx; red[8 + wid] = mxx; red[16 + wid] = mny; red[24 + wid] = mxy; }
        __syncthreads();
        mnx = red[0]; mxx = red[8]; mny = red[16]; mxy = red[24];
        for (int i = 1; i < nw; ++i) {
            mnx = fmin(mnx, red[i]); mxx = fmax(mxx, red[8 + i]); mny = fmin(mny, red[16 + i]); mxy = fmax(mxy, red[24 + i]);
        }
    }
    // pixel axes (:91, 95-106)
    double T0, Tl, U0, Ul;
    if (g.has_fpgrid) {
        T0 = __ddiv_rn(__dsub_rn(g.fp_t0, g.t0), delt);
        Tl = __ddiv_rn(__dsub_rn(g.fp_t1, g.t0), delt);
        U0 = __ddiv_rn(__dsub_rn(g.fp_u0, u0), du);
        Ul = __ddiv_rn(__dsub_rn(g.fp_u1, u0), du);
    } else {
        T0 = o.pn[0].x; Tl = o.pn[nt - 1].x; U0 = 0.0; Ul = 1.0;
    }
    const double Ts = ntg > 1 ? __ddiv_rn(__dsub_rn(Tl, T0), (double)(ntg - 1)) : 0.0;
    const double Us = nug > 1 ? __ddiv_rn(__dsub_rn(Ul, U0), (double)(nug - 1)) : 0.0;
    // FP32 frame: origin at the centre of the bounding box of samples and pixels,
    // scaled by a power of two so that every distance in the box is < 1.
    const double bx0 = fmin(mnx, fmin(T0, Tl)), bx1 = fmax(mxx, fmax(T0, Tl));
    const double by0 = fmin(mny, fmin(U0, Ul)), by1 = fmax(mxy, fmax(U0, Ul));
    const double ccx = 0.5 * (bx0 + bx1), ccy = 0.5 * (by0 + by1);
    const double diag = sqrt((bx1 - bx0) * (bx1 - bx0) + (by1 - by0) * (by1 - by0));
    int ex = 0;
    frexp(diag > 0.0 ? diag : 1.0, &ex);           // diag = m * 2^ex, m in [0.5, 1)
    const double sigma = ldexp(1.0, -ex);
    const int S = nt - 1;
    const int Spad = ((S + kTilePad - 1) / kTilePad) * kTilePad;
    int degen = 0;
    float* const tabE = reinterpret_cast<float*>(o.A);
    float* const tabM = reinterpret_cast<float*>(o.A + (Spad >> 1));
    for (int s = tid; s < Spad; s += nth) {
        float4 A;
        float h;
        if (s < S) {
            const double2 a = o.pn[s], b = o.pn[s + 1];
            const double cx = b.x - a.x, cy = b.y - a.y;
            const double len = sqrt(cx * cx + cy * cy);
            double exd = 1.0, eyd = 0.0;
            if (len > 0.0) { exd = cx / len; eyd = cy / len; } else { degen++; }
            const double mx = (a.x + 0.5 * cx - ccx) * sigma, my = (a.y + 0.5 * cy - ccy) * sigma;
            const float fex = (float)exd, fey = (float)eyd;
            const float am = (float)(mx * exd + my * eyd);     // mid . e
            const float bm = (float)(mx * eyd - my * exd);     // perp' = px*ey - py*ex - bm
            A = make_float4(fex, fey, -am, -bm);
            h = (float)(0.5 * len * sigma);
        } else {   // padding: along = -4 -> sat -> 1, perp = 2 -> D = kPadD
            A = make_float4(0.f, 0.f, -4.f, 2.f);
            h = 0.f;
        }
        const int at = (s >> 1) * 4 + (s & 1);
        tabE[at] = A.x; tabE[at + 2] = A.y; tabM[at] = A.z; tabM[at + 2] = A.w;
        o.H[s] = h;
    }
    for (int tile = tid; o.bbox != nullptr && tile < Spad / o.tile; tile += nth) {
        const int s0 = min(tile * o.tile, S), s1 = min(s0 + o.tile, S);      // vertices s0 .. s1 inclusive
        double xlo = CUDART_INF, xhi = -CUDART_INF, ylo = CUDART_INF, yhi = -CUDART_INF;
        for (int j = s0; j <= s1; ++j) {
            const double2 p = o.pn[j];
            xlo = fmin(xlo, p.x); xhi = fmax(xhi, p.x); ylo = fmin(ylo, p.y); yhi = fmax(yhi, p.y);
        }
        o.bbox[tile] = make_float4((float)((xlo - ccx) * sigma), (float)((xhi - ccx) * sigma),
                                   (float)((ylo - ccy) * sigma), (float)((yhi - ccy) * sigma));
    }
    for (int i = tid; i < ntg; i += nth)
        o.pxs[i] = (float)((lin_axis(T0, Ts, Tl, i, ntg) - ccx) * sigma);
    for (int i = tid; i < nug; i += nth)
        o.pys[i] = (float)((lin_axis(U0, Us, Ul, i, nug) - ccy) * sigma);
    if (degen) atomicAdd(&o.hdr->degenerate, degen);   // zeroed by the caller before prep_window
    if (tid == 0) {
        WinHdr* h = o.hdr;
        h->T0 = T0; h->Tstep = Ts; h->Tlast = Tl; h->U0 = U0; h->Ustep = Us; h->Ulast = Ul;
        h->ccx = ccx; h->ccy = ccy; h->sigma = sigma; h->du = du; h->u0raw = u0raw; h->u1raw = u1raw;
    }
}

// ------------------------------------------------------------------ per-pixel epilogue values
struct PixelVals {
    double d, pdf, xcx, xcy, g;   // g = (xc_y - p_y)/d   (dddx_y, libs/FingerprintLib.py:355)
};

__device__ __forceinline__ PixelVals pixel_values(const double2* __restrict__ pn, const PixelHit& hit,
                                                  double py, double lambda, int q) {
    PixelVals v;
    const double2 a = pn[hit.s], b = pn[hit.s + 1];
    const double cx = __dsub_rn(b.x, a.x), cy = __dsub_rn(b.y, a.y);
    v.xcx = __dadd_rn(a.x, __dmul_rn(hit.lam, cx));            // xclose (:262)
    v.xcy = __dadd_rn(a.y, __dmul_rn(hit.lam, cy));
    v.d = __dsqrt_rn(hit.D);                                   // (:263)
    v.pdf = (q == 2) ? exp(-(v.d * v.d) / lambda) : exp(-fabs(v.d) / lambda);   // (:174,176)
    v.g = (v.xcy - py) / v.d;
    return v;
}

}  // namespace wfot
