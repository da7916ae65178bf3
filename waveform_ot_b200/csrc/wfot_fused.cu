// wfot_fused.cu -- the throughput path: one persistent kernel that takes a
// waveform window from raw samples to (W^t, W^u, dW^t/dw, dW^u/dw, dW^t/dx0)
// without the 2-D field ever leaving the chip's caches as an HBM-sized array.
//
// One CTA owns one window at a time (grid = resident CTAs, windows strided):
//   P0  prep_window            FP64 normalisation + FP32 segment table -> shared memory
//   P1  scan_block + resolve   nearest segment per pixel (FP32 scan with exact tile pruning over
//                              warp footprints drawn from a shared counter, FP64 exact tie
//                              resolution); per pixel {iray, pdf, wa, wb} go to a per-CTA
//                              scratch slab (28 B/pixel, re-used for every window of the CTA)
//   P2  marginals              fixed-order column / row sums of pdf -> time / amplitude
//                              marginals of the normalised density (OTlib.py:92-93,155-156)
//   P3  block_ot1d x 2         CDF scan, merge, W_p^p, dW/df, dW/dx0 per marginal
//                              (OTlib.py:596-706) and <dW, pbar> (OTlib.py:1141,1144-1145)
//   P4  gradient assembly      sum_k pdf_k (R_k - Rbar)/A dd_k/dw_j keyed by iray
//                              (FingerprintLib.py:205-228), run-combined per pixel column,
//                              accumulated with FP64 reductions in L2
// Reference chain replaced: ricker_util.py:386-388 (BuildOTobjfromWaveform ->
// CalcWasserWaveform(deriv=True, returnmarg=True)).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "wfot_device.cuh"
#include "wfot_host.h"
#include "wfot_ot.cuh"

namespace wfot {

constexpr int kFQCap = 512;
struct FQEntry { int pix; float b1; };


// Shared-memory layout (byte offsets from the dynamic shared base).  Pointers are formed
// from the `extern __shared__` symbol inside each kernel so the compiler keeps them in the
// shared address space (LDS/STS instead of generic loads).
struct SmemLayout {
    int pn, A, H, bbox, keys, pxs, pys, margt, margu, Rt, Ru, xt, xu, cf, E, tk, dx, red, gbins, posf, queue, hdr, qcount;
    int total;
};

inline SmemLayout make_layout(int nt, int Spad, int ntg_pad, int nug_pad, int nmax) {
    SmemLayout L;
    int o = 0;
    auto take = [&](int bytes) { const int at = o; o += (bytes + 15) & ~15; return at; };
    L.pn = take(nt * 16);
    // union region: the FP32 segment table is only needed by the scan / resolve phase (P0-P1);
    // the OT scratch (P3) and the per-sample chain factors (P4) re-use its bytes.
    const int ubase = o;
    L.A = take(Spad * 16);
    L.H = take(Spad * 4);
    const int ntiles = Spad / tile_for(nt);
    L.bbox = take(ntiles * 16);
    L.keys = take(8 * ntiles * 4);                // best-first tile keys, one array per warp (<= 8 warps)
    const int uend_scan = o;
    o = ubase;
    L.cf = take(nmax * 8);
    L.E = take(nmax * 8);
    L.tk = take(nmax * 16);
    L.dx = take(nmax * 16);
    L.posf = take(nmax * 4);
    L.gbins = take(nt * 8);
    o = o > uend_scan ? o : uend_scan;
    L.margt = take(ntg_pad * 8);
    L.margu = take(nug_pad * 8);
    L.Rt = take(ntg_pad * 8);
    L.Ru = take(nug_pad * 8);
    L.xt = take(ntg_pad * 8);
    L.xu = take(nug_pad * 8);
    L.red = take(64 * 8);
    L.hdr = take(128);
    L.queue = take(kFQCap * (int)sizeof(FQEntry));
    L.pxs = take(ntg_pad * 4);
    L.pys = take(nug_pad * 4);
    L.qcount = take(16);
    L.total = o;
    return L;
}

struct FusedArgs {
    const void* t; const void* w; int dtype; long long t_stride; int nt;
    const wfot_grid* grids; int n_grids; int B; int nug, ntg;
    double lambda; int q, pmask, transform;
    const double* tgt_cdf_t; const double* tgt_x_t; const double* tgt_cdf_u; const double* tgt_x_u;
    int tgt_rows;
    double* W; double* grad; double* dwg;
    // per-CTA scratch slabs
    double* s_pdf; double* s_wa; double* s_wb; int32_t* s_idx;
    int32_t* status;
    int* next_window;     // global work counter (zeroed by the launcher)
    int cluster;          // > 1: launched as thread-block clusters of this many CTAs, one window per CLUSTER
    int Spad, ntg_pad, nug_pad, nmax;
    SmemLayout L;
};

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
    int* const s_posf = reinterpret_cast<int*>(smem_raw + (L).posf);                     \
    FQEntry* const s_queue = reinterpret_cast<FQEntry*>(smem_raw + (L).queue);           \
    WinHdr* const s_hdr = reinterpret_cast<WinHdr*>(smem_raw + (L).hdr);                 \
    int* const s_qcount = reinterpret_cast<int*>(smem_raw + (L).qcount)

// pixel -> scratch: density and the two gradient weights of its nearest segment
__device__ __forceinline__ void store_pixel(const FusedArgs& a, const double2* pn, size_t slab,
                                            int it, int iu, const PixelHit& hit, double py, int& zero_dist) {
    const PixelVals v = pixel_values(pn, hit, py, a.lambda, a.q);
    const size_t k = slab + (size_t)iu * a.ntg + it;
    double wgt = v.pdf * v.g;                                  // pdf * dddx_y (FingerprintLib.py:355)
    if (a.q == 2) wgt *= 2.0 * fabs(v.d);                      // :214-217
    zero_dist += (v.d == 0.0);
    a.s_pdf[k] = v.pdf;
    a.s_wa[k] = (1.0 - hit.lam) * wgt;                         // -> sample iray    (:223)
    a.s_wb[k] = hit.lam * wgt;                                 // -> sample iray+1  (:224)
    a.s_idx[k] = hit.s;
}

// NT threads per CTA: 256 for large grids; 128 (4 CTAs per SM) for small windows, where the per-window
// phases are short and more co-resident windows hide the block barriers between them.
template <int R, int NT, int T>
__global__ void __launch_bounds__(NT, 512 / NT) k_misfit_grad(FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WFOT_SMEM_POINTERS(a.L);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npix = a.nug * a.ntg, S = a.nt - 1;
    // Small batches (fewer windows than resident CTAs) are launched as thread-block clusters: the CTAs of a
    // cluster share ONE window - each prepares it in its own shared memory, they split the pixel footprints
    // of P1, and rank 0 runs the short P2-P4 after a cluster barrier.  That cuts the latency of a single
    // evaluation (the reference's scipy.optimize loops evaluate one model at a time).
    namespace cg = cooperative_groups;
    const int csize = a.cluster > 1 ? a.cluster : 1;
    const int crank = csize > 1 ? (int)cg::this_cluster().block_rank() : 0;
    const int cid = (int)blockIdx.x / csize, nclusters = (int)gridDim.x / csize;
    const size_t slab = (size_t)cid * npix;
    SegTable tb{s_A, s_H, s_bbox, S, a.Spad, T};
    int zero_dist = 0, slow = 0, common = 0, degen = 0, tiles = 0;

    // windows are drawn from a global counter (the pruned scan makes their cost uneven): the first window of a
    // CTA is blockIdx.x, the following ones come from the counter, which starts at gridDim.x.  Clusters take
    // windows cid, cid + nclusters, ... (every CTA of a cluster must see the same sequence).
    for (int b = cid; b < a.B;) {
        // ---------------- P0: window -> shared memory
        const wfot_grid g = a.grids[b % a.n_grids];
        if (tid == 0) {
            s_hdr->degenerate = 0; s_qcount[0] = 0; s_qcount[1] = 0;
            s_qcount[2] = csize > 1 ? b + nclusters
                                    : (int)gridDim.x + atomicAdd(a.next_window, 1);      // this CTA's next window
        }
        if (a.grad && crank == 0) {      // P4 accumulates into these rows with L2 reductions
            double* const g0 = a.grad + ((size_t)b * 2) * a.nt;
            for (int j = tid; j < 2 * a.nt; j += NT) g0[j] = 0.0;
            __threadfence();
        }
        __syncthreads();
        PrepOut po{s_pn, s_A, s_H, s_bbox, tb.tile, s_pxs, s_pys, s_hdr};
        prep_window(a.t, a.w, a.dtype, (long long)b * a.t_stride, (long long)b * a.nt, a.nt, g,
                    a.nug, a.ntg, a.transform, po, s_red, nullptr);
        __syncthreads();
        const WinHdr hdr = *s_hdr;
        degen += (tid == 0 && crank == 0) ? hdr.degenerate : 0;
        for (int i = tid; i < a.ntg; i += NT) s_xt[i] = lin_axis(hdr.T0, hdr.Tstep, hdr.Tlast, i, a.ntg);
        for (int i = tid; i < a.nug; i += NT) s_xu[i] = lin_axis(hdr.U0, hdr.Ustep, hdr.Ulast, i, a.nug);
        __syncthreads();

        // ---------------- P1: nearest segment per pixel.  Warps draw footprints from a shared counter
        //                  (the pruned scan makes their cost uneven).
        const FootMap fm = make_footmap<R>(a.ntg, a.nug, fabsf(s_pxs[a.ntg - 1] - s_pxs[0]),
                                           fabsf(s_pys[a.nug - 1] - s_pys[0]));
        for (;;) {
            int f = 0;
            if (lane == 0) f = atomicAdd(s_qcount + 1, 1) * csize + crank;     // this CTA's share of the footprints
            f = __shfl_sync(0xffffffffu, f, 0);
            if (f >= fm.nfoot) break;
            const LaneBlock lb = lane_block<R>(fm, f, lane, a.ntg, a.nug, s_pxs, s_pys);
            const int cp = lb.cp, rg = lb.rg;
            const int it0 = 2 * cp, it1 = min(2 * cp + 1, a.ntg - 1);
            float py[R];
#pragma unroll
            for (int r = 0; r < R; ++r) py[r] = s_pys[min(rg * R + r, a.nug - 1)];
            float b1[2 * R], b2[2 * R], b3[2 * R];
            int t1[2 * R];
            scan_block<R, T>(tb, lb.fp, s_pxs[it0], s_pxs[it1], py, b1, t1, b2, b3, tiles,
                             s_keys + (threadIdx.x >> 5) * (a.Spad / T));
            if (!lb.owns) continue;
            float lb1[2 * R], lb2[2 * R], lb3[2 * R];
            int lt1[2 * R];
#pragma unroll
            for (int k = 0; k < 2 * R; ++k) { lb1[k] = b1[k]; lb2[k] = b2[k]; lb3[k] = b3[k]; lt1[k] = t1[k]; }
#pragma unroll 1
            for (int k = 0; k < 2 * R; ++k) {
                const int it = 2 * cp + (k & 1), iu = rg * R + (k >> 1);
                if (it >= a.ntg || iu >= a.nug) continue;
                const float kb1 = lb1[k];
                const double pyd = s_xu[iu];
                PixelHit hit;
                if (!resolve_pixel<T>(tb, s_pn, s_pxs[it], s_pys[iu], s_xt[it], pyd, kb1, lt1[k], lb2[k], lb3[k], hit)) {
                    const int qi = atomicAdd(s_qcount, 1);
                    if (qi < kFQCap) { s_queue[qi] = FQEntry{iu * a.ntg + it, kb1}; continue; }
                    ++slow;
                    resolve_pixel_full(tb, s_pn, s_pxs[it], s_pys[iu], s_xt[it], pyd, kb1, hit);
                }
                store_pixel(a, s_pn, slab, it, iu, hit, pyd, zero_dist);
            }
        }
        __syncthreads();
        {
            const int nq = min(*s_qcount, kFQCap);
            for (int e = warp; e < nq; e += NT / 32) {
                const FQEntry qe = s_queue[e];
                const int it = qe.pix % a.ntg, iu = qe.pix / a.ntg;
                PixelHit hit;
                resolve_pixel_warp(tb, s_pn, s_pxs[it], s_pys[iu], s_xt[it], s_xu[iu], qe.b1, hit);
                if (lane == 0) { store_pixel(a, s_pn, slab, it, iu, hit, s_xu[iu], zero_dist); ++slow; }
            }
        }
        __syncthreads();   // scratch slab complete (block-scope visibility of global writes)
        if (csize > 1) {   // ... and cluster-scope: every CTA's share of the slab is visible to rank 0
            __threadfence();
            cg::this_cluster().sync();
        }
        if (crank == 0) {
        // ---------------- P2: marginals of the normalised density (fixed summation order;
        //                  8 independent loads in flight per thread)
        double part = 0.0;
        for (int c = tid; c < a.ntg; c += NT) {
            const double* col = a.s_pdf + slab + c;
            double s0 = 0.0;
            int iu = 0;
            for (; iu + 8 <= a.nug; iu += 8) {
                double v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __ldcg(col + (size_t)(iu + j) * a.ntg);
#pragma unroll
                for (int j = 0; j < 8; ++j) s0 += v[j];
            }
            for (; iu < a.nug; ++iu) s0 += __ldcg(col + (size_t)iu * a.ntg);
            s_margt[c] = s0;
            part += s0;
        }
        const double A = block_sum(part, s_red);                        // OTpdf.amp (OTlib.py:92)
        for (int iu = warp; iu < a.nug; iu += NT / 32) {
            const double* row = a.s_pdf + slab + (size_t)iu * a.ntg;
            double s0 = 0.0;
            for (int c = lane; c < a.ntg; c += 32) s0 += __ldcg(row + c);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, off);
            if (lane == 0) s_margu[iu] = s0 / A;                        // OTlib.py:93,156
        }
        for (int c = tid; c < a.ntg; c += NT) s_margt[c] = s_margt[c] / A;     // OTlib.py:93,155
        __syncthreads();

        // ---------------- P3: 1-D OT per marginal
        const size_t trow = (size_t)(b % a.tgt_rows);
        OtScratch sc{s_cf, s_tk, s_dx, s_E, s_posf, s_red};
        for (int c = tid; c < a.ntg; c += NT) s_cf[c] = s_margt[c];
        __syncthreads();
        const OtResult rt = block_ot1d(sc, a.ntg, a.tgt_cdf_t + trow * a.ntg, a.ntg, s_xt,
                                       a.tgt_x_t + trow * a.ntg, a.pmask,
                                       (a.pmask & 1) ? s_Rt : nullptr, (a.pmask & 2) ? s_Rt : nullptr, nullptr);
        double gp = 0.0;
        for (int c = tid; c < a.ntg; c += NT) gp += s_margt[c] * s_Rt[c];
        const double Gt = block_sum(gp, s_red);                         // <dwpmargX, pbar> (OTlib.py:1144)
        for (int c = tid; c < a.nug; c += NT) s_cf[c] = s_margu[c];
        __syncthreads();
        const OtResult ru = block_ot1d(sc, a.nug, a.tgt_cdf_u + trow * a.nug, a.nug, s_xu,
                                       a.tgt_x_u + trow * a.nug, a.pmask,
                                       (a.pmask & 1) ? s_Ru : nullptr, (a.pmask & 2) ? s_Ru : nullptr, nullptr);
        gp = 0.0;
        for (int c = tid; c < a.nug; c += NT) gp += s_margu[c] * s_Ru[c];
        const double Gu = block_sum(gp, s_red);                         // OTlib.py:1145
        common += (tid == 0) ? (rt.common + ru.common) : 0;
        if (tid == 0) {
            a.W[2 * (size_t)b] = (a.pmask & 1) ? rt.W1 : rt.W2;
            a.W[2 * (size_t)b + 1] = (a.pmask & 1) ? ru.W1 : ru.W2;
            if (a.dwg) a.dwg[b] = (a.pmask & 1) ? rt.dpos1 : rt.dpos2;  // OTlib.py:1121
        }

        // ---------------- P4: gradient assembly (FingerprintLib.py:205-228).  A thread walks a pixel
        //                  column, combines runs of equal nearest segment and adds each run to the
        //                  window's gradient rows with fire-and-forget FP64 reductions in L2
        //                  (RED.ADD.F64; shared-memory FP64 atomics are CAS loops).  The rows were
        //                  zeroed in P0.
        if (a.grad) {
            // chain vectors: (R - <R, pbar>)/A  (OTlib.py:1144-1147)
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
            for (int c = tid; c < a.ntg; c += NT) {
                const double ct = s_Rt[c];
                int cur = -1;
                double t0 = 0.0, t1 = 0.0, u0 = 0.0, u1 = 0.0;
                for (int iu0 = 0; iu0 < a.nug; iu0 += 4) {
                    int idx[4];
                    double wa[4], wb[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const size_t k = slab + (size_t)min(iu0 + j, a.nug - 1) * a.ntg + c;
                        idx[j] = __ldcg(a.s_idx + k); wa[j] = __ldcg(a.s_wa + k); wb[j] = __ldcg(a.s_wb + k);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (iu0 + j >= a.nug) break;
                        if (idx[j] != cur) {
                            if (cur >= 0) {
                                const double c0 = s_gbins[cur], c1 = s_gbins[cur + 1];
                                atomicAdd(gt + cur, t0 * c0); atomicAdd(gt + cur + 1, t1 * c1);
                                atomicAdd(gu + cur, u0 * c0); atomicAdd(gu + cur + 1, u1 * c1);
                            }
                            cur = idx[j]; t0 = t1 = u0 = u1 = 0.0;
                        }
                        const double cu = s_Ru[iu0 + j];
                        t0 += wa[j] * ct; t1 += wb[j] * ct; u0 += wa[j] * cu; u1 += wb[j] * cu;
                    }
                }
                if (cur >= 0) {
                    const double c0 = s_gbins[cur], c1 = s_gbins[cur + 1];
                    atomicAdd(gt + cur, t0 * c0); atomicAdd(gt + cur + 1, t1 * c1);
                    atomicAdd(gu + cur, u0 * c0); atomicAdd(gu + cur + 1, u1 * c1);
                }
            }
        }
        }   // crank == 0
        b = s_qcount[2];
        __syncthreads();
        if (csize > 1) cg::this_cluster().sync();   // rank 0 is done reading the slab: the next window may overwrite it
    }
    if (a.status) {
        if (zero_dist) atomicAdd(a.status + WFOT_STAT_ZERO_DIST, zero_dist);
        if (slow) atomicAdd(a.status + WFOT_STAT_SLOW_PIXELS, slow);
        if (common) atomicAdd(a.status + WFOT_STAT_COMMON_CDF, common);
        if (degen) atomicAdd(a.status + WFOT_STAT_DEGENERATE_SEG, degen);
        if (lane == 0 && tiles)
            atomicAdd(reinterpret_cast<unsigned long long*>(a.status + WFOT_STAT_SCAN_TILES),
                      (unsigned long long)tiles * (R / 4) * (T / 8));
    }
}

// ------------------------------------------------------------------ scan-only probe
// prep_window + scan_block, nothing else: FP32 nearest distance per pixel (no FP64 resolve).
// Used by bench.py to attribute time between the brute-force scan and the epilogues.
template <int R>
__global__ void __launch_bounds__(256, 2) k_scan_probe(FusedArgs a, float* out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WFOT_SMEM_POINTERS(a.L);
    const int tid = threadIdx.x, S = a.nt - 1;
    SegTable tb{s_A, s_H, s_bbox, S, a.Spad, 16};
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const wfot_grid g = a.grids[b % a.n_grids];
        if (tid == 0) s_hdr->degenerate = 0;
        __syncthreads();
        PrepOut po{s_pn, s_A, s_H, s_bbox, tb.tile, s_pxs, s_pys, s_hdr};
        prep_window(a.t, a.w, a.dtype, (long long)b * a.t_stride, (long long)b * a.nt, a.nt, g,
                    a.nug, a.ntg, 0, po, s_red, nullptr);
        __syncthreads();
        const float inv_sigma = (float)(1.0 / s_hdr->sigma);
        const FootMap fm = make_footmap<R>(a.ntg, a.nug, fabsf(s_pxs[a.ntg - 1] - s_pxs[0]),
                                           fabsf(s_pys[a.nug - 1] - s_pys[0]));
        int tiles = 0;
        for (int f = tid >> 5; f < fm.nfoot; f += 8) {
            const LaneBlock lb = lane_block<R>(fm, f, tid & 31, a.ntg, a.nug, s_pxs, s_pys);
            const int cp = lb.cp, rg = lb.rg;
            const int it0 = 2 * cp, it1 = min(2 * cp + 1, a.ntg - 1);
            float py[R];
#pragma unroll
            for (int r = 0; r < R; ++r) py[r] = s_pys[min(rg * R + r, a.nug - 1)];
            float b1[2 * R], b2[2 * R], b3[2 * R];
            int t1[2 * R];
            scan_block<R, 16>(tb, lb.fp, s_pxs[it0], s_pxs[it1], py, b1, t1, b2, b3, tiles,
                              s_keys + (threadIdx.x >> 5) * (a.Spad / 16));
            if (!lb.owns) continue;
#pragma unroll
            for (int k = 0; k < 2 * R; ++k) {
                const int it = 2 * cp + (k & 1), iu = rg * R + (k >> 1);
                if (it < a.ntg && iu < a.nug)
                    out[((size_t)b * a.nug + iu) * a.ntg + it] = sqrtf(b1[k]) * inv_sigma + 0.f * (b2[k] + b3[k] + t1[k]);
            }
        }
        __syncthreads();
        (void)s_margt; (void)s_margu; (void)s_Rt; (void)s_Ru; (void)s_xt; (void)s_xu; (void)s_cf; (void)s_E;
        (void)s_tk; (void)s_dx; (void)s_gbins; (void)s_posf; (void)s_queue; (void)s_qcount;
    }
}

// Rows per thread-owned pixel block (2 columns x R rows).  R = 8 halves the per-segment
// set-up work per pixel; R = 4 halves the register footprint (3 CTAs per SM) and gives 16 x 16
// pixel warp footprints, which prune ~20 % more segment tiles.  WFOT_DEV_R overrides (tuning aid).
static int rows_per_thread() {
    const char* e = getenv("WFOT_DEV_R");
    if (e && e[0] == '4') return 4;
    if (e && e[0] == '8') return 8;
    return 4;
}

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

}  // namespace wfot

using namespace wfot;

extern "C" {

// Scratch: 28 bytes per pixel per resident CTA.  Sized for the largest grid the
// device can co-schedule (SM count x 8 CTAs) so the query needs no device call.
size_t wfot_misfit_grad_workspace_bytes(int B, int nt, int nug, int ntg) {
    if (B <= 0 || nt < 2 || nug < 1 || ntg < 1) return 0;
    int sms = wfot_device_sm_count();
    if (sms <= 0) sms = 148;
    size_t ctas = (size_t)sms * 8;
    if ((size_t)B < ctas) ctas = (size_t)B;
    return ctas * (size_t)nug * ntg * 28 + 512;
}

int wfot_misfit_grad_batch(const void* t, const void* w, int in_dtype, long long t_stride, int nt,
                           const wfot_grid* grids, int n_grids, int B, int nug, int ntg, double lambda,
                           int q, int pmask, int transform, const double* tgt_cdf_t, const double* tgt_x_t,
                           const double* tgt_cdf_u, const double* tgt_x_u, int tgt_rows, double* W,
                           double* grad, double* dwg, void* workspace, size_t workspace_bytes,
                           int32_t* status, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!t || !w || !grids || !W || !workspace || !tgt_cdf_t || !tgt_x_t || !tgt_cdf_u || !tgt_x_u ||
        B <= 0 || nt < 2 || nug < 1 || ntg < 1 || n_grids < 1 || tgt_rows < 1 ||
        (q != 0 && q != 2) || (pmask != WFOT_W1 && pmask != WFOT_W2) || !(lambda > 0.0) ||
        (in_dtype != WFOT_F32 && in_dtype != WFOT_F64))
        return WFOT_ERR_INVALID_ARG;
    FusedArgs a;
    a.t = t; a.w = w; a.dtype = in_dtype; a.t_stride = t_stride; a.nt = nt; a.grids = grids;
    a.n_grids = n_grids; a.B = B; a.nug = nug; a.ntg = ntg; a.lambda = lambda; a.q = q; a.pmask = pmask;
    a.transform = transform; a.tgt_cdf_t = tgt_cdf_t; a.tgt_x_t = tgt_x_t; a.tgt_cdf_u = tgt_cdf_u;
    a.tgt_x_u = tgt_x_u; a.tgt_rows = tgt_rows; a.W = W; a.grad = grad; a.dwg = dwg;
    a.status = status;
    a.Spad = seg_pad(nt); a.ntg_pad = pad4(ntg); a.nug_pad = pad4(nug);
    a.nmax = pad4(ntg > nug ? ntg : nug);
    a.L = make_layout(nt, a.Spad, a.ntg_pad, a.nug_pad, a.nmax);
    const size_t smem = (size_t)a.L.total;
    if (smem > 227 * 1024) return WFOT_ERR_UNSUPPORTED;
    int per_sm = 0;
    const int R = rows_per_thread();
    // small windows: 128-thread CTAs (4 per SM), tiny ones 64-thread CTAs (8 per SM)
    const char* ent = getenv("WFOT_DEV_NT");
    int nthreads = 256;
    if (R == 4 && (long long)nug * ntg <= 16384) nthreads = 128;
    if (R == 4 && (long long)nug * ntg <= 8192) nthreads = 64;
    if (R == 4 && ent) nthreads = atoi(ent) == 64 ? 64 : atoi(ent) == 128 ? 128 : 256;
    const int T = tile_for(nt);
    int ctas;
#define WFOT_OCC(RR, NN, TT) resident_ctas(k_misfit_grad<RR, NN, TT>, smem, &per_sm, NN)
    if (R == 8) ctas = WFOT_OCC(8, 256, 16);
    else if (T == 8) ctas = nthreads == 64 ? WFOT_OCC(4, 64, 8) : nthreads == 128 ? WFOT_OCC(4, 128, 8) : WFOT_OCC(4, 256, 8);
    else ctas = nthreads == 64 ? WFOT_OCC(4, 64, 16) : nthreads == 128 ? WFOT_OCC(4, 128, 16) : WFOT_OCC(4, 256, 16);
#undef WFOT_OCC
    if (ctas < 1) return cuda_fail(cudaGetLastError(), "k_misfit_grad occupancy");
    // few windows: clusters of 2/4/8 CTAs per window (as many as keep every window resident at once)
    int csize = 1;
    {
        const char* ec = getenv("WFOT_DEV_CLUSTER");
        const int cmax = ec ? atoi(ec) : 8;
        while (csize * 2 <= cmax && (long long)B * csize * 2 <= ctas) csize *= 2;
    }
    a.cluster = csize;
    if (csize > 1) ctas = B * csize;
    else if (ctas > B) ctas = B;
    const size_t npix = (size_t)nug * ntg;
    uintptr_t base = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    if (workspace_bytes < (base - (uintptr_t)workspace) + 256) return WFOT_ERR_WORKSPACE;
    a.next_window = (int*)base;                       // first 256 bytes: the window counter
    if (cudaMemsetAsync(a.next_window, 0, 256, stream) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaMemsetAsync");
    base += 256;
    const size_t avail = workspace_bytes - (base - (uintptr_t)workspace);
    const size_t max_ctas = avail / (npix * 28);
    if (max_ctas < 1) return WFOT_ERR_WORKSPACE;
    if (csize == 1 && (size_t)ctas > max_ctas) ctas = (int)max_ctas;
    if (csize > 1 && (size_t)(ctas / csize) > max_ctas) return WFOT_ERR_WORKSPACE;   // one slab per cluster
    unsigned char* p = (unsigned char*)base;
    const size_t nslab = csize > 1 ? (size_t)(ctas / csize) : (size_t)ctas;
    a.s_pdf = (double*)p;   p += nslab * npix * 8;
    a.s_wa = (double*)p;    p += nslab * npix * 8;
    a.s_wb = (double*)p;    p += nslab * npix * 8;
    a.s_idx = (int32_t*)p;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3((unsigned)ctas); cfg.blockDim = dim3((unsigned)nthreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    if (csize > 1) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
    }
    cudaError_t le = cudaSuccess;
#define WFOT_LAUNCH(RR, NN, TT) le = cudaLaunchKernelEx(&cfg, k_misfit_grad<RR, NN, TT>, a)
    if (R == 8) WFOT_LAUNCH(8, 256, 16);
    else if (T == 8) { if (nthreads == 64) WFOT_LAUNCH(4, 64, 8); else if (nthreads == 128) WFOT_LAUNCH(4, 128, 8); else WFOT_LAUNCH(4, 256, 8); }
    else { if (nthreads == 64) WFOT_LAUNCH(4, 64, 16); else if (nthreads == 128) WFOT_LAUNCH(4, 128, 16); else WFOT_LAUNCH(4, 256, 16); }
#undef WFOT_LAUNCH
    if (le != cudaSuccess) return cuda_fail(le, "wfot_misfit_grad_batch launch");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "wfot_misfit_grad_batch launch");
    return WFOT_OK;
}

int wfot_scan_probe(const void* t, const void* w, int in_dtype, long long t_stride, int nt,
                    const wfot_grid* grids, int n_grids, int B, int nug, int ntg, float* dist32,
                    void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!t || !w || !grids || !dist32 || B <= 0 || nt < 2 || nug < 1 || ntg < 1) return WFOT_ERR_INVALID_ARG;
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.t = t; a.w = w; a.dtype = in_dtype; a.t_stride = t_stride; a.nt = nt; a.grids = grids;
    a.n_grids = n_grids; a.B = B; a.nug = nug; a.ntg = ntg;
    a.Spad = seg_pad(nt); a.ntg_pad = pad4(ntg); a.nug_pad = pad4(nug);
    a.nmax = pad4(ntg > nug ? ntg : nug);
    a.L = make_layout(nt, a.Spad, a.ntg_pad, a.nug_pad, a.nmax);
    const size_t smem = (size_t)a.L.total;
    if (smem > 227 * 1024) return WFOT_ERR_UNSUPPORTED;
    int per_sm = 0;
    const int R = rows_per_thread();
    int ctas = (R == 4) ? resident_ctas(k_scan_probe<4>, smem, &per_sm) : resident_ctas(k_scan_probe<8>, smem, &per_sm);
    if (ctas < 1) return cuda_fail(cudaGetLastError(), "k_scan_probe occupancy");
    if (ctas > B) ctas = B;
    if (R == 4) k_scan_probe<4><<<ctas, 256, smem, stream>>>(a, dist32);
    else k_scan_probe<8><<<ctas, 256, smem, stream>>>(a, dist32);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "wfot_scan_probe launch");
    return WFOT_OK;
}

}  // extern "C"
