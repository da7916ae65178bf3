// wfot_fused.cuh -- pieces shared by the two forms of the throughput path:
//   wfot_fused.cu   k_misfit_grad   one persistent kernel per call (small batches: thread-block
//                                   clusters share a window; short windows)
//   wfot_split.cu   k_scan + k_resolve   the same work as two kernels for large batches of large
//                                   windows: the FP32 scan keeps its 120+ registers, the FP64
//                                   resolve / density / OT / gradient half runs at twice the occupancy
// Shared: argument block, shared-memory layout, the per-pixel scratch-slab store and the window
// tail (marginals -> 1-D OT -> gradient assembly, phases P2-P4).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "wfot_device.cuh"
#include "wfot_exp.cuh"
#include "wfot_host.h"
#include "wfot_ot.cuh"

namespace wfot {

constexpr int kFQCap = 512;
struct FQEntry { int pix; float b1; };
// Column sums of the density are defined as: rows split into kRowGroups groups by (row mod kRowGroups), each group
// summed in ascending row order, the group sums combined as ((g0+g1)+(g2+g3))+((g4+g5)+(g6+g7)).  Every form of the
// path (slab pass of the single-kernel form, warp-per-row accumulation of the resolve kernel) realises this order,
// so the marginals - and the CDFs compared for exact equality - do not depend on the launch shape.
constexpr int kRowGroups = 8;

// Shared-memory layout (byte offsets from the dynamic shared base).  Pointers are formed
// from the `extern __shared__` symbol inside each kernel so the compiler keeps them in the
// shared address space (LDS/STS instead of generic loads).
struct SmemLayout {
    int pn, A, H, bbox, keys, pxs, pys, margt, margu, Rt, Ru, xt, xu, cf, E, tk, dx, red, gbins, posf, queue, hdr, qcount;
    int cf2, E2, tk2, dx2, posf2;   // OT scratch of the amplitude marginal (0: absent, the two problems share one warp)
    int etab;       // 2^(j/64) table of exp_neg() (1 KB)
    int colpart;    // resolve kernel: per-row-group column sums of the density, [kRowGroups][ntg_pad] doubles
    int qcap;       // entries of the ambiguous-pixel queue
    int total;
};

enum LayoutKind { kLayoutFused = 0, kLayoutScan = 1, kLayoutResolve = 2 };

// kLayoutScan keeps only what prep_window + scan_block touch; kLayoutResolve drops the tile boxes and keys (and
// halves the queue) so that four 256-thread CTAs fit one SM.
inline SmemLayout make_layout(int nt, int Spad, int ntg_pad, int nug_pad, int nmax, int kind = kLayoutFused,
                              bool pair = false) {
    SmemLayout L;
    memset(&L, 0, sizeof(L));
    int o = 0;
    auto take = [&](int bytes) { const int at = o; o += (bytes + 15) & ~15; return at; };
    L.pn = take(nt * 16);
    // union region: the FP32 segment table is only needed by the scan / resolve phase (P0-P1);
    // the OT scratch (P3) and the per-sample chain factors (P4) re-use its bytes.
    const int ubase = o;
    L.A = take(Spad * 16);
    L.H = take(Spad * 4);
    if (kind != kLayoutResolve) {
        const int ntiles = Spad / tile_for(nt);
        L.bbox = take(ntiles * 16);
        L.keys = take(8 * ntiles * 4);            // best-first tile keys, one array per warp (<= 8 warps)
    }
    const int uend_scan = o;
    if (kind != kLayoutScan) {
        o = ubase;
        L.cf = take(nmax * 8);
        L.E = take(nmax * 8);
        L.tk = take(nmax * 16);
        L.dx = take(nmax * 16);
        L.posf = take(nmax * 4);
        if (pair) {      // the two marginal problems side by side, one warp each (warp_ot1d)
            L.cf2 = take(nmax * 8);
            L.E2 = take(nmax * 8);
            L.tk2 = take(nmax * 16);
            L.dx2 = take(nmax * 16);
            L.posf2 = take(nmax * 4);
        }
        L.gbins = take(nt * 8);
        o = o > uend_scan ? o : uend_scan;
        L.margt = take(ntg_pad * 8);
        L.margu = take(nug_pad * 8);
        L.Rt = take(ntg_pad * 8);
        L.Ru = take(nug_pad * 8);
        L.xt = take(ntg_pad * 8);
        L.xu = take(nug_pad * 8);
        if (kind == kLayoutResolve) L.colpart = take(kRowGroups * ntg_pad * 8);
        L.qcap = (kind == kLayoutResolve) ? kFQCap / 2 : kFQCap;
        L.queue = take(L.qcap * (int)sizeof(FQEntry));
    }
    if (kind != kLayoutScan) L.etab = take(128 * 8);
    L.red = take(64 * 8);
    L.hdr = take(128);
    L.pxs = take(ntg_pad * 4);
    L.pys = take(nug_pad * 4);
    L.qcount = take(16);
    L.total = o;
    return L;
}

// Layout with the second OT scratch set when that does not cost a resident CTA (`cta_cap` = CTAs per SM the registers allow)
inline SmemLayout make_layout_auto(int nt, int Spad, int ntg_pad, int nug_pad, int nmax, int kind, int cta_cap) {
    const SmemLayout L0 = make_layout(nt, Spad, ntg_pad, nug_pad, nmax, kind, false);
    const SmemLayout L1 = make_layout(nt, Spad, ntg_pad, nug_pad, nmax, kind, true);
    auto ctas = [&](int total) { const int c = (227 * 1024) / (total + 1024); return c < cta_cap ? c : cta_cap; };
    return (L1.total <= 227 * 1024 && ctas(L1.total) >= ctas(L0.total)) ? L1 : L0;
}

struct FusedArgs {
    const void* t; const void* w; int dtype; long long t_stride; int nt;
    const wfot_grid* grids; int n_grids; int B; int nug, ntg;
    double lambda, rlambda; int q, pmask, transform;      // rlambda = RN(1 / lambda)
    const double* tgt_cdf_t; const double* tgt_x_t; const double* tgt_cdf_u; const double* tgt_x_u;
    int tgt_rows;
    double* W; double* grad; double* dwg;
    // observed-window mode (wfot_marginal_cdfs_batch): stop after the CDFs of the two marginals and write them
    // (B, ntg) / (B, nug) and the 2-D amplitude (B,) instead of running the OT against a target
    double* out_cdf_t; double* out_cdf_u; double* out_amp;
    // measurement aid (wfot_dev.h): nearest-segment index of every pixel of every window, (B, nug * ntg) int32
    int32_t* dbg_iray;
    unsigned long long* dbg_phase;   // measurement aid (wfot_dev.h): per-phase cycle counters of k_resolve
    // per-CTA scratch slabs
    // s_w: one 16-byte entry per pixel, {weight of sample idx, weight of sample idx + 1 with idx in its 16 lowest
    // mantissa bits (pack_wbi)}
    double* s_pdf; ulonglong2* s_w;
    int32_t* status;
    int* next_window;     // global work counters (zeroed by the launcher): [0] fused / scan, [16] resolve
    int cluster;          // > 1: launched as thread-block clusters of this many CTAs, one window per CLUSTER
    // split form: windows b0 .. b0 + B - 1 of the call; per-pixel scan results (tile | flags, 2 bytes) of window
    // b0 + i at scan_out[i * npix ...]
    int b0;
    uint16_t* scan_out;
    int Spad, ntg_pad, nug_pad, nmax;
    SmemLayout L;
};

// flags in the second word of a scan result: another tile / two other tiles hold a segment within the FP32
// rounding tolerance of the pixel's minimum (see resolve_pixel)
constexpr unsigned kScanFlag2 = 1u << 14, kScanFlag3 = 1u << 15, kScanTileMask = kScanFlag2 - 1u;   // 16-bit results: < 16384 tiles

#define WFOT_SMEM_POINTERS(L)                                                            \
    double2* const s_pn = reinterpret_cast<double2*>(smem_raw + (L).pn);                 \
    float4* const s_A = reinterpret_cast<float4*>(smem_raw + (L).A);                     \
    float* const s_H = reinterpret_cast<float*>(smem_raw + (L).H);                       \
    float4* const s_bbox = reinterpret_cast<float4*>(smem_raw + (L).bbox);               \
    unsigned* const s_keys = reinterpret_cast<unsigned*>(smem_raw + (L).keys);           \
    float* const s_pxs = reinterpret_cast<float*>(smem_raw + (L).pxs);                   \
    float* const s_pys = reinterpret_cast<float*>(smem_raw + (L).pys);                   \
    double* const s_margt = reinterpret_cast<double*>(smem_raw + (L).margt);             \
    double* const s_margu = reinterpret_cast<double*>(smem_raw + (L).margu);             \
    double* const s_Rt = reinterpret_cast<double*>(smem_raw + (L).Rt);                   \
    double* const s_Ru = reinterpret_cast<double*>(smem_raw + (L).Ru);                   \
    double* const s_xt = reinterpret_cast<double*>(smem_raw + (L).xt);                   \
    double* const s_xu = reinterpret_cast<double*>(smem_raw + (L).xu);                   \
    double* const s_cf = reinterpret_cast<double*>(smem_raw + (L).cf);                   \
    double* const s_E = reinterpret_cast<double*>(smem_raw + (L).E);                     \
    double* const s_tk = reinterpret_cast<double*>(smem_raw + (L).tk);                   \
    double* const s_dx = reinterpret_cast<double*>(smem_raw + (L).dx);                   \
    double* const s_red = reinterpret_cast<double*>(smem_raw + (L).red);                 \
    double* const s_gbins = reinterpret_cast<double*>(smem_raw + (L).gbins);             \
    double* const s_colpart = reinterpret_cast<double*>(smem_raw + (L).colpart);         \
    double2* const s_etab = reinterpret_cast<double2*>(smem_raw + (L).etab);             \
    int* const s_posf = reinterpret_cast<int*>(smem_raw + (L).posf);                     \
    FQEntry* const s_queue = reinterpret_cast<FQEntry*>(smem_raw + (L).queue);           \
    WinHdr* const s_hdr = reinterpret_cast<WinHdr*>(smem_raw + (L).hdr);                 \
    int* const s_qcount = reinterpret_cast<int*>(smem_raw + (L).qcount)

// pixel -> scratch: density and the two gradient weights of its nearest segment.
// The throughput path needs d = sqrt(D) only to form exp(-d / lambda) and (xclose_y - p_y) / d, so the three
// slow FP64 library sequences (IEEE sqrt, two IEEE divisions: ~70 instructions) are replaced by one reciprocal
// square root (rsqrt_lean) and fused-multiply-add corrections: d = D r + (D - (D r)^2) r / 2, the quotients by Markstein's q0 = a y,
// q = q0 + (a - b q0) y with y = RN(1 / b) (correctly rounded but for rare half-ulp cases; the density then
// differs from the materialising kernel's by <= 1 ulp of the exponent, i.e. ~1e-15 relative).
// 1 / sqrt(D) for normal D > 0: one MUFU (rsqrt.approx.f64, ~2^-22) and a third-order correction
// y = y0 (1 + e/2 + 3 e^2 / 8), e = 1 - D y0^2 (truncation 5 e^3 / 16 ~ 2^-66): 6 instructions against ~17 for rsqrt().
__device__ __forceinline__ double rsqrt_lean(double D) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(D));
    const double e = fma(-(D * y0), y0, 1.0);
    return fma(y0 * e, fma(0.375, e, 0.5), y0);
}

// Copy the exp table to shared memory (every thread of the CTA calls; a barrier must follow before store_pixel).
__device__ __forceinline__ void load_exp_table(double2* s_etab) {
    for (int i = threadIdx.x; i < 128; i += blockDim.x) reinterpret_cast<double*>(s_etab)[i] = kExp2Tab[i];
}

// Slab entry of a pixel: 16 bytes.  wa (weight of sample idx) as a double; wb (weight of sample idx + 1) rounded to
// a 36-bit mantissa (relative error 2^-37: the gradient keeps ~11 digits) with idx in the 16 bits that frees.
// Every byte of the slab crosses DRAM twice (the CTAs of a launch write far more than L2 holds) and the resolve
// kernel's gradient phase runs at the speed of that read-back.
__device__ __forceinline__ unsigned long long pack_wbi(double wb, int idx) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(wb) + 0x8000ull;    // round to nearest
    return (u & ~0xffffull) | (unsigned long long)(unsigned)idx;
}
__device__ __forceinline__ void unpack_wbi(unsigned long long u, double& wb, int& idx) {
    idx = (int)(u & 0xffffull);
    wb = __longlong_as_double((long long)(u & ~0xffffull));
}

template <bool STORE_PDF = true>
__device__ __forceinline__ double store_pixel(const FusedArgs& a, const double2* pn, const double2* etab, size_t slab,
                                              int it, int iu, const PixelHit& hit, double py, int& zero_dist,
                                              int32_t* dbg_iray = nullptr, bool live = true) {
    const double* const pny = reinterpret_cast<const double*>(pn) + 1;
    const double ay = pny[2 * hit.s], by = pny[2 * hit.s + 2];
    const double cy = __dsub_rn(by, ay);
    const double xcy = __dadd_rn(ay, __dmul_rn(hit.lam, cy));  // xclose_y (FingerprintLib.py:262)
    const double num = xcy - py;
    double d, g;
    if (hit.D > 0.0) {
        const double rs = rsqrt_lean(hit.D);
        const double d0 = hit.D * rs;
        d = fma(fma(-d0, d0, hit.D), 0.5 * rs, d0);            // (:263)
        const double g0 = num * rs;
        g = fma(fma(-g0, d, num), rs, g0);                     // dddx_y = (xclose_y - p_y) / d (:355)
    } else {                                                   // pixel on the waveform: 0/0 as in the reference
        d = 0.0;
        g = num / d;
        zero_dist += live ? 1 : 0;
    }
    const double e = (a.q == 2) ? d * d : d;                   // exp(-d^2/lambda) (:174) or exp(-|d|/lambda) (:176)
    const double q0 = e * a.rlambda;
    const double x = fma(fma(-q0, a.lambda, e), a.rlambda, q0);
    const double pdf = exp_neg(x, etab);
    double wgt = pdf * g;                                      // pdf * dddx_y
    if (a.q == 2) wgt *= 2.0 * d;                              // :214-217
    const size_t k = slab + (size_t)iu * a.ntg + it;
    // weights of samples iray / iray + 1 (:223-224).  A pixel whose nearest point is the far vertex of its segment
    // (lam = 1: all of its weight goes to sample iray + 1) is filed under the next segment with lam = 0 - the same
    // contribution - so that the pixels around a vertex form one run for the gradient assembly (P4).
    const bool far = hit.lam == 1.0 && hit.s + 2 < a.nt && (wgt - wgt == 0.0);   // finite weight: NaN goes where the reference puts it
    const double wa = far ? wgt : (1.0 - hit.lam) * wgt, wb = far ? 0.0 : hit.lam * wgt;
    if (live) {                        // a shadow lane (row tail) computes and stores nothing
        if (STORE_PDF) a.s_pdf[k] = pdf;
        a.s_w[k] = make_ulonglong2((unsigned long long)__double_as_longlong(wa), pack_wbi(wb, hit.s + (far ? 1 : 0)));
        if (dbg_iray) dbg_iray[(size_t)iu * a.ntg + it] = hit.s;
    }
    return pdf;
}

// ------------------------------------------------------------------ window tail: P2-P4
// Called by every thread of the CTA that owns window b once its scratch slab is complete and visible.
//   P2  marginals of the normalised density (OTlib.py:92-93,155-156), fixed summation order
//   P3  1-D OT per marginal (OTlib.py:596-706) and <dW, pbar> (OTlib.py:1141,1144-1145)
//   P4  gradient assembly (FingerprintLib.py:205-228): a thread walks a pixel column, combines runs of
//       equal nearest segment and adds each run to the window's gradient rows with fire-and-forget FP64
//       reductions in L2 (RED.ADD.F64; shared-memory FP64 atomics are CAS loops).  The rows were zeroed in P0.
// Returns the number of exact source/target CDF coincidences (libs/OTlib.py:663-666) in thread 0.
template <int NT, bool HAVE_SUMS = false, int P4R = 4>
__device__ __forceinline__ int window_tail(const FusedArgs& a, unsigned char* smem_raw, int b, size_t slab,
                                           const WinHdr& hdr, long long* t_p4 = nullptr) {
    WFOT_SMEM_POINTERS(a.L);
    (void)s_pn; (void)s_A; (void)s_H; (void)s_bbox; (void)s_keys; (void)s_pxs; (void)s_pys; (void)s_queue;
    (void)s_hdr; (void)s_qcount; (void)s_colpart; (void)s_etab;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---------------- P2: column sums (kRowGroups groups of rows, see above)
    if (HAVE_SUMS) {   // the resolve kernel accumulated the groups (s_colpart) and the raw row sums (s_margu) in P1
        for (int c = tid; c < a.ntg; c += NT) {
            const double* g = s_colpart + c;
            const int st = a.ntg_pad;
            s_margt[c] = ((g[0] + g[st]) + (g[2 * st] + g[3 * st])) + ((g[4 * st] + g[5 * st]) + (g[6 * st] + g[7 * st]));
        }
    } else {
        for (int c = tid; c < a.ntg; c += NT) {
            const double* col = a.s_pdf + slab + c;
            double g[kRowGroups];
#pragma unroll
            for (int j = 0; j < kRowGroups; ++j) g[j] = 0.0;
            int iu = 0;
            for (; iu + kRowGroups <= a.nug; iu += kRowGroups) {
                double v[kRowGroups];
#pragma unroll
                for (int j = 0; j < kRowGroups; ++j) v[j] = __ldcg(col + (size_t)(iu + j) * a.ntg);
#pragma unroll
                for (int j = 0; j < kRowGroups; ++j) g[j] += v[j];
            }
#pragma unroll
            for (int j = 0; j < kRowGroups; ++j)
                if (iu + j < a.nug) g[j] += __ldcg(col + (size_t)(iu + j) * a.ntg);
            s_margt[c] = ((g[0] + g[1]) + (g[2] + g[3])) + ((g[4] + g[5]) + (g[6] + g[7]));
        }
    }
    __syncthreads();
    const double A = canon_sum(s_margt, a.ntg, s_red);               // OTpdf.amp (OTlib.py:92), launch-shape independent
    if (HAVE_SUMS) {
        for (int iu = tid; iu < a.nug; iu += NT) s_margu[iu] = s_margu[iu] / A;
    } else {
        // row sums: lane l adds columns l, l + 32, ... in ascending order, then a xor-shuffle tree
        for (int iu = warp; iu < a.nug; iu += NT / 32) {
            const double* row = a.s_pdf + slab + (size_t)iu * a.ntg;
            double s0 = 0.0;
            for (int c = lane; c < a.ntg; c += 32) s0 += __ldcg(row + c);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, off);
            if (lane == 0) s_margu[iu] = s0 / A;                        // OTlib.py:93,156
        }
    }
    for (int c = tid; c < a.ntg; c += NT) s_margt[c] = s_margt[c] / A;     // OTlib.py:93,155
    __syncthreads();

    if (a.out_cdf_t) {   // observed-window mode: OTpdf of the two marginals (OTlib.py:157-160 -> :91-93,112-114)
        for (int c = tid; c < a.ntg; c += NT) s_cf[c] = s_margt[c];
        __syncthreads();
        int neg = 0;
        canon_cdf(s_cf, a.ntg, s_red, &neg);
        for (int c = tid; c < a.ntg; c += NT) a.out_cdf_t[(size_t)b * a.ntg + c] = s_cf[c];
        __syncthreads();
        for (int c = tid; c < a.nug; c += NT) s_cf[c] = s_margu[c];
        __syncthreads();
        canon_cdf(s_cf, a.nug, s_red, &neg);
        for (int c = tid; c < a.nug; c += NT) a.out_cdf_u[(size_t)b * a.nug + c] = s_cf[c];
        if (tid == 0 && a.out_amp) a.out_amp[b] = A;
        __syncthreads();
        return 0;
    }

    // ---------------- P3: both marginal problems by warp 0, without block barriers (warp_ot1d)
    const size_t trow = (size_t)(b % a.tgt_rows);
    OtScratch sc{s_cf, s_tk, s_dx, s_E, s_posf, s_red};
    const int dmask = a.grad ? a.pmask : 0;      // amplitude derivatives of the 1-D problems only when a gradient is assembled
    for (int c = tid; c < a.ntg; c += NT) s_cf[c] = s_margt[c];
    __syncthreads();
    if (a.L.cf2) {       // one warp per marginal
        double* const s_cf2 = reinterpret_cast<double*>(smem_raw + a.L.cf2);
        if (warp == 0) {
            const WarpOtResult ot = warp_ot1d(sc, a.ntg, a.tgt_cdf_t + trow * a.ntg, s_xt, a.tgt_x_t + trow * a.ntg,
                                              dmask, s_Rt, s_margt);
            if (lane == 0) {
                s_red[0] = ot.r.W1; s_red[1] = ot.r.W2; s_red[2] = ot.r.dpos1; s_red[3] = ot.r.dpos2; s_red[4] = ot.G;
                s_red[8] = __longlong_as_double((long long)ot.r.common);
            }
        } else if (warp == 1) {
            const OtScratch sc2{s_cf2, reinterpret_cast<double*>(smem_raw + a.L.tk2),
                                reinterpret_cast<double*>(smem_raw + a.L.dx2), reinterpret_cast<double*>(smem_raw + a.L.E2),
                                reinterpret_cast<int*>(smem_raw + a.L.posf2), s_red};
            for (int c = lane; c < a.nug; c += 32) s_cf2[c] = s_margu[c];
            __syncwarp();
            const WarpOtResult ou = warp_ot1d(sc2, a.nug, a.tgt_cdf_u + trow * a.nug, s_xu, a.tgt_x_u + trow * a.nug,
                                              dmask, s_Ru, s_margu);
            if (lane == 0) {
                s_red[5] = ou.r.W1; s_red[6] = ou.r.W2; s_red[7] = ou.G;
                s_red[9] = __longlong_as_double((long long)ou.r.common);
            }
        }
    } else if (warp == 0) {
        const WarpOtResult ot = warp_ot1d(sc, a.ntg, a.tgt_cdf_t + trow * a.ntg, s_xt, a.tgt_x_t + trow * a.ntg,
                                          dmask, s_Rt, s_margt);
        __syncwarp();
        for (int c = lane; c < a.nug; c += 32) s_cf[c] = s_margu[c];
        __syncwarp();
        const WarpOtResult ou = warp_ot1d(sc, a.nug, a.tgt_cdf_u + trow * a.nug, s_xu, a.tgt_x_u + trow * a.nug,
                                          dmask, s_Ru, s_margu);
        if (lane == 0) {
            s_red[0] = ot.r.W1; s_red[1] = ot.r.W2; s_red[2] = ot.r.dpos1; s_red[3] = ot.r.dpos2; s_red[4] = ot.G;
            s_red[5] = ou.r.W1; s_red[6] = ou.r.W2; s_red[7] = ou.G;
            s_red[8] = __longlong_as_double((long long)ot.r.common);
            s_red[9] = __longlong_as_double((long long)ou.r.common);
        }
    }
    __syncthreads();
    OtResult rt, ru;
    rt.W1 = s_red[0]; rt.W2 = s_red[1]; rt.dpos1 = s_red[2]; rt.dpos2 = s_red[3];
    ru.W1 = s_red[5]; ru.W2 = s_red[6];
    const double Gt = s_red[4], Gu = s_red[7];                      // <dwpmarg, pbar> (OTlib.py:1144-1145)
    rt.common = (int)__double_as_longlong(s_red[8]); ru.common = (int)__double_as_longlong(s_red[9]);
    __syncthreads();                                                // s_red is re-used below
    if (tid == 0) {
        if (a.pmask == 3) {   // both orders from one fingerprint (misfit only): [W1^t, W1^u, W2^t, W2^u], dwg [W1, W2]
            double* const w4 = a.W + 4 * (size_t)b;
            w4[0] = rt.W1; w4[1] = ru.W1; w4[2] = rt.W2; w4[3] = ru.W2;
            if (a.dwg) { a.dwg[2 * (size_t)b] = rt.dpos1; a.dwg[2 * (size_t)b + 1] = rt.dpos2; }
        } else {
            a.W[2 * (size_t)b] = (a.pmask & 1) ? rt.W1 : rt.W2;
            a.W[2 * (size_t)b + 1] = (a.pmask & 1) ? ru.W1 : ru.W2;
            if (a.dwg) a.dwg[b] = (a.pmask & 1) ? rt.dpos1 : rt.dpos2;  // OTlib.py:1121
        }
    }

    // ---------------- P4
    if (t_p4) *t_p4 = clock64();
    if (a.grad) {
        // chain vectors: (R - <R, pbar>)/A  (OTlib.py:1144-1147)
        // the window's gradient rows are zeroed here, not at the start of the window: the L2 reductions below then
        // find their lines in L2 (a window's worth of slab traffic would have evicted them in between)
        {
            double* const g0 = a.grad + ((size_t)b * 2) * a.nt;
            for (int j = tid; j < 2 * a.nt; j += NT) g0[j] = 0.0;
        }
        for (int c = tid; c < a.ntg; c += NT) s_Rt[c] = (s_Rt[c] - Gt) / A;
        for (int c = tid; c < a.nug; c += NT) s_Ru[c] = (s_Ru[c] - Gu) / A;
        const double scale = -1.0 / (a.lambda * hdr.du);             // FingerprintLib.py:228,376-378
        for (int j = tid; j < a.nt; j += NT) {
            double chain = scale;
            if (a.transform) {   // d(un)/du, ricker_util.py:273,393-397
                const double wj = load_sample(a.w, a.dtype, (long long)b * a.nt + j);
                const double up = ((wj - hdr.u0raw) + (wj - hdr.u1raw)) / (hdr.u1raw - hdr.u0raw);
                chain *= 2.0 / ((hdr.u1raw - hdr.u0raw) * CUDART_PI * (1.0 + up * up));
            }
            s_gbins[j] = chain;
        }
        __syncthreads();
        double* const gt = a.grad + ((size_t)b * 2) * a.nt;
        double* const gu = gt + a.nt;
        // one L2 reduction per (sample, row of the gradient); nothing is sent for a sum of exact zeros (the weights of
        // the 60-80 % of pixels whose nearest point is a vertex)
        auto flush = [&](int j, double tv, double uv) {
            if (tv != 0.0 || uv != 0.0) {
                const double cj = s_gbins[j];
                red_add(gt + j, tv * cj); red_add(gu + j, uv * cj);
            }
        };
        // a thread walks a column; narrow grids are cut into row bands so that every warp has a column to walk
        // (79 x 61 pixels: 4 bands of 20 rows x 64 lanes instead of 61 busy threads out of 256)
        const int cpp = min(NT, (a.ntg + 31) & ~31), nbands = NT / cpp;
        const int rows_band = (a.nug + nbands - 1) / nbands;
        const int r_lo = min((tid / cpp) * rows_band, a.nug), r_hi = min(r_lo + rows_band, a.nug);
        for (int c = tid % cpp; c < a.ntg && r_lo < r_hi; c += cpp) {
            const double ct = s_Rt[c];
            // (t0, u0) / (t1, u1): running sums for samples cur / cur + 1.  When the nearest segment moves to a
            // neighbouring one, the sums of the sample the two segments share are carried over instead of being sent.
            int cur = -2;
            double t0 = 0.0, t1 = 0.0, u0 = 0.0, u1 = 0.0;
            for (int iu0 = r_lo; iu0 < r_hi; iu0 += P4R) {
                int idx[P4R];
                double wa[P4R], wb[P4R];
                unsigned long long wbi[P4R];
#pragma unroll
                for (int j = 0; j < P4R; ++j) {      // P4R rows in flight (the slab comes back from DRAM)
                    const size_t k = slab + (size_t)min(iu0 + j, r_hi - 1) * a.ntg + c;
                    const ulonglong2 e = __ldcg(a.s_w + k);
                    wa[j] = __longlong_as_double((long long)e.x); wbi[j] = e.y;
                }
#pragma unroll
                for (int j = 0; j < P4R; ++j) unpack_wbi(wbi[j], wb[j], idx[j]);
#pragma unroll
                for (int j = 0; j < P4R; ++j) {
                    if (iu0 + j >= r_hi) break;
                    const int d = idx[j] - cur;
                    if (d != 0) {      // one body for the three cases (lanes of a warp hit different ones in the same row)
                        const bool up = d == 1, dn = d == -1;
                        if (!dn) flush(cur, t0, u0);              // sample cur is complete unless the run moved down
                        if (!up) flush(cur + 1, t1, u1);          // sample cur + 1 is complete unless it moved up
                        const double c0 = up ? t1 : 0.0, c1 = up ? u1 : 0.0, c2 = dn ? t0 : 0.0, c3 = dn ? u0 : 0.0;
                        t0 = c0; u0 = c1; t1 = c2; u1 = c3;
                        cur = idx[j];
                    }
                    const double cu = s_Ru[iu0 + j];
                    t0 += wa[j] * ct; t1 += wb[j] * ct; u0 += wa[j] * cu; u1 += wb[j] * cu;
                }
            }
            if (cur >= 0) { flush(cur, t0, u0); flush(cur + 1, t1, u1); }
        }
    }
    return (tid == 0) ? (rt.common + ru.common) : 0;
}

// resident CTAs of a kernel at a given dynamic shared-memory size (sets the opt-in limit on the way)
template <typename K>
static int resident_ctas(K kernel, size_t smem, int* per_sm_out, int threads = 256) {
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess) return -1;
    if (per_sm < 1) return -1;
    if (per_sm_out) *per_sm_out = per_sm;
    return sms * per_sm;
}

// wfot_split.cu: the two-kernel form.  `a` comes fully populated except scan_out / b0 / the slab pointers;
// `ws` / `ws_bytes` = what is left of the caller's workspace after the 256-byte counter block at a.next_window.
int launch_split(FusedArgs a, unsigned char* ws, size_t ws_bytes, cudaStream_t stream);
size_t split_workspace_bytes(int B, int nt, int nug, int ntg, int sms);
// true when the two-kernel form is used for this problem size (large batches of large windows)
bool split_wanted(int B, int nt, int nug, int ntg, int sms);

}  // namespace wfot
