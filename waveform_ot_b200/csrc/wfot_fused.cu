// wfot_fused.cu -- the throughput path: one persistent kernel that takes a
// waveform window from raw samples to (W^t, W^u, dW^t/dw, dW^u/dw, dW^t/dx0)
// without the 2-D field ever leaving the chip's caches as an HBM-sized array.
//
// One CTA owns one window at a time (grid = resident CTAs, windows drawn from a counter):
//   P0  prep_window            FP64 normalisation + FP32 segment table -> shared memory
//   P1  scan_block + resolve   nearest segment per pixel (FP32 scan with exact tile pruning over
//                              warp footprints drawn from a shared counter, FP64 exact tie
//                              resolution); per pixel {iray, pdf, wa, wb} go to a per-CTA
//                              scratch slab (24 B/pixel, re-used for every window of the CTA)
//   P2-P4  window_tail()       marginals, 1-D OT per marginal, gradient assembly (wfot_fused.cuh)
// Large batches of large windows take the two-kernel form of the same phases (wfot_split.cu).
// Reference chain replaced: ricker_util.py:386-388 (BuildOTobjfromWaveform ->
// CalcWasserWaveform(deriv=True, returnmarg=True)).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/wfot_dev.h"
#include "wfot_fused.cuh"
#include "wfot_dev_options.h"

namespace wfot {

// NT threads per CTA: 256 for large grids; 128 (4 CTAs per SM) for small windows, where the per-window
// phases are short and more co-resident windows hide the block barriers between them.
template <int R, int NT, int T>
__global__ void __launch_bounds__(NT, 512 / NT) k_misfit_grad(FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WFOT_SMEM_POINTERS(a.L);
    (void)s_margt; (void)s_margu; (void)s_Rt; (void)s_Ru; (void)s_cf; (void)s_E; (void)s_tk; (void)s_dx;
    (void)s_gbins; (void)s_posf; (void)s_colpart;
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
    load_exp_table(s_etab);      // visible behind the first barrier of the window loop

    // windows are drawn from a global counter (the pruned scan makes their cost uneven): the first window of a
    // CTA is blockIdx.x, the following ones come from the counter, which starts at gridDim.x.  Clusters take
    // windows cid, cid + nclusters, ... (every CTA of a cluster must see the same sequence).
    for (int b = cid; b < a.B;) {
        // ---------------- P0: window -> shared memory
        const wfot_grid g = a.grids[b % a.n_grids];
        if (tid == 0) { s_hdr->degenerate = 0; s_qcount[0] = 0; s_qcount[1] = 0; }
        __syncthreads();
        // this CTA's next window: asked for now, needed after the tail (the round trip to L2 hides behind P1)
        if (tid == 0) s_qcount[2] = csize > 1 ? b + nclusters : (int)gridDim.x + atomicAdd(a.next_window, 1);
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
        int32_t* const dbg = a.dbg_iray ? a.dbg_iray + (size_t)b * npix : nullptr;
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
                    if (qi < a.L.qcap) { s_queue[qi] = FQEntry{iu * a.ntg + it, kb1}; continue; }
                    ++slow;
                    resolve_pixel_full(tb, s_pn, s_pxs[it], s_pys[iu], s_xt[it], pyd, kb1, hit);
                }
                store_pixel(a, s_pn, s_etab, slab, it, iu, hit, pyd, zero_dist, dbg);
            }
        }
        __syncthreads();
        {
            const int nq = min(*s_qcount, a.L.qcap);
            for (int e = warp; e < nq; e += NT / 32) {
                const FQEntry qe = s_queue[e];
                const int it = qe.pix % a.ntg, iu = qe.pix / a.ntg;
                PixelHit hit;
                resolve_pixel_warp(tb, s_pn, s_pxs[it], s_pys[iu], s_xt[it], s_xu[iu], qe.b1, hit);
                if (lane == 0) { store_pixel(a, s_pn, s_etab, slab, it, iu, hit, s_xu[iu], zero_dist, dbg); ++slow; }
            }
        }
        __syncthreads();   // scratch slab complete (block-scope visibility of global writes)
        if (csize > 1) {   // ... and cluster-scope: every CTA's share of the slab is visible to rank 0
            __threadfence();
            cg::this_cluster().sync();
        }
        if (crank == 0) {
            common += window_tail<NT>(a, smem_raw, b, slab, hdr);
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

}  // namespace wfot

namespace wfot {
// wfot_dev_epilogue_math: the fused path's exp(-x) and 1/sqrt(x) on arbitrary arguments (accuracy tests)
__global__ void __launch_bounds__(256) k_epilogue_math(const double* __restrict__ x, double* __restrict__ e,
                                                       double* __restrict__ r, int n) {
    __shared__ double2 tab[64];
    load_exp_table(tab);
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) { e[i] = exp_neg(x[i], tab); r[i] = rsqrt_lean(x[i]); }
}
}  // namespace wfot

using namespace wfot;

namespace wfot {
static int g_dev_options[kOptCount] = {0};
static int32_t* g_iray_capture = nullptr;
static unsigned long long* g_phase_cycles = nullptr;
int dev_option(int id) { return (id >= 0 && id < kOptCount) ? g_dev_options[id] : 0; }
}  // namespace wfot

// Shared launcher: `a` holds the problem, outputs and mode; picks the form (one kernel / scan + resolve),
// the CTA shape and the cluster size, carves the workspace and launches.
static int run_fused(FusedArgs a, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int nt = a.nt, nug = a.nug, ntg = a.ntg, B = a.B;
    a.rlambda = 1.0 / a.lambda;
    a.dbg_iray = g_iray_capture;
    a.dbg_phase = g_phase_cycles;
    a.Spad = seg_pad(nt); a.ntg_pad = pad4(ntg); a.nug_pad = pad4(nug);
    a.nmax = pad4(ntg > nug ? ntg : nug);
    a.L = make_layout_auto(nt, a.Spad, a.ntg_pad, a.nug_pad, a.nmax, kLayoutFused,
                           (long long)nug * ntg <= 8192 ? 8 : (long long)nug * ntg <= 16384 ? 4 : 2);
    const size_t smem = (size_t)a.L.total;
    if (smem > 227 * 1024 || nt > 65536) return WFOT_ERR_UNSUPPORTED;      // (slab entries carry 16-bit segment indices)
    uintptr_t base = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    if (workspace_bytes < (base - (uintptr_t)workspace) + 256) return WFOT_ERR_WORKSPACE;
    a.next_window = (int*)base;                       // first 256 bytes: the window counters
    if (cudaMemsetAsync(a.next_window, 0, 256, stream) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaMemsetAsync");
    base += 256;
    const size_t avail = workspace_bytes - (base - (uintptr_t)workspace);
    int sms = wfot_device_sm_count();
    if (sms <= 0) return cuda_fail(cudaGetLastError(), "device query");
    const int pipeline = dev_option(kOptPipeline);
    if (pipeline != 1 && split_wanted(B, nt, nug, ntg, sms))
        return launch_split(a, (unsigned char*)base, avail, stream);

    int per_sm = 0;
    // small windows: 128-thread CTAs (4 per SM), tiny ones 64-thread CTAs (8 per SM)
    int nthreads = 256;
    if ((long long)nug * ntg <= 16384) nthreads = 128;
    if ((long long)nug * ntg <= 8192) nthreads = 64;
    if (const int ont = dev_option(kOptFusedThreads)) nthreads = ont == 64 ? 64 : ont == 128 ? 128 : 256;
    const int T = tile_for(nt);
    int ctas;
#define WFOT_OCC(NN, TT) resident_ctas(k_misfit_grad<4, NN, TT>, smem, &per_sm, NN)
    if (T == 8) ctas = nthreads == 64 ? WFOT_OCC(64, 8) : nthreads == 128 ? WFOT_OCC(128, 8) : WFOT_OCC(256, 8);
    else ctas = nthreads == 64 ? WFOT_OCC(64, 16) : nthreads == 128 ? WFOT_OCC(128, 16) : WFOT_OCC(256, 16);
#undef WFOT_OCC
    if (ctas < 1) return cuda_fail(cudaGetLastError(), "k_misfit_grad occupancy");
    // few windows: clusters of 2/4/8 CTAs per window (as many as keep every window resident at once)
    int csize = 1;
    {
        const int oc = dev_option(kOptClusterMax);
        const int cmax = oc > 0 ? oc : 8;
        while (csize * 2 <= cmax && (long long)B * csize * 2 <= ctas) csize *= 2;
    }
    a.cluster = csize;
    if (csize > 1) ctas = B * csize;
    else if (ctas > B) ctas = B;
    const size_t npix = (size_t)nug * ntg;
    const size_t max_ctas = avail / (npix * 24);
    if (max_ctas < 1) return WFOT_ERR_WORKSPACE;
    if (csize == 1 && (size_t)ctas > max_ctas) ctas = (int)max_ctas;
    if (csize > 1 && (size_t)(ctas / csize) > max_ctas) return WFOT_ERR_WORKSPACE;   // one slab per cluster
    unsigned char* p = (unsigned char*)base;
    const size_t nslab = csize > 1 ? (size_t)(ctas / csize) : (size_t)ctas;
    a.s_pdf = (double*)p;   p += nslab * npix * 8;
    p = (unsigned char*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
    a.s_w = (ulonglong2*)p;
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
#define WFOT_LAUNCH(NN, TT) le = cudaLaunchKernelEx(&cfg, k_misfit_grad<4, NN, TT>, a)
    if (T == 8) { if (nthreads == 64) WFOT_LAUNCH(64, 8); else if (nthreads == 128) WFOT_LAUNCH(128, 8); else WFOT_LAUNCH(256, 8); }
    else { if (nthreads == 64) WFOT_LAUNCH(64, 16); else if (nthreads == 128) WFOT_LAUNCH(128, 16); else WFOT_LAUNCH(256, 16); }
#undef WFOT_LAUNCH
    if (le != cudaSuccess) return cuda_fail(le, "wfot_misfit_grad_batch launch");
    note_launches(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "wfot_misfit_grad_batch launch");
    return WFOT_OK;
}

extern "C" {

int wfot_dev_set_option(int id, int value) {
    if (id < 0 || id >= kOptCount) return -1;
    const int old = g_dev_options[id];
    g_dev_options[id] = value;
    return old;
}

// Scratch: 28 bytes per pixel per resident CTA (24 used; sized for SM count x 8 CTAs so the query needs no
// occupancy call), plus - for batches that take the two-kernel form - 2 bytes per pixel per window
// of one scan/resolve launch pair.
void wfot_dev_capture_iray(int32_t* iray) { g_iray_capture = iray; }
void wfot_dev_phase_cycles(unsigned long long* cycles) { g_phase_cycles = cycles; }

int wfot_dev_epilogue_math(const double* x, double* exp_neg_out, double* rsqrt_out, int n, void* stream) {
    if (!x || !exp_neg_out || !rsqrt_out || n <= 0) return WFOT_ERR_INVALID_ARG;
    k_epilogue_math<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, exp_neg_out, rsqrt_out, n);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? WFOT_OK : cuda_fail(e, "wfot_dev_epilogue_math launch");
}

size_t wfot_misfit_grad_workspace_bytes(int B, int nt, int nug, int ntg) {
    if (B <= 0 || nt < 2 || nug < 1 || ntg < 1) return 0;
    int sms = wfot_device_sm_count();
    if (sms <= 0) sms = 148;
    size_t ctas = (size_t)sms * 8;
    if ((size_t)B < ctas) ctas = (size_t)B;
    return ctas * (size_t)nug * ntg * 28 + 512 + split_workspace_bytes(B, nt, nug, ntg, sms);
}

int wfot_misfit_grad_batch(const void* t, const void* w, int in_dtype, long long t_stride, int nt,
                           const wfot_grid* grids, int n_grids, int B, int nug, int ntg, double lambda,
                           int q, int pmask, int transform, const double* tgt_cdf_t, const double* tgt_x_t,
                           const double* tgt_cdf_u, const double* tgt_x_u, int tgt_rows, double* W,
                           double* grad, double* dwg, void* workspace, size_t workspace_bytes,
                           int32_t* status, void* stream_) {
    if (!t || !w || !grids || !W || !workspace || !tgt_cdf_t || !tgt_x_t || !tgt_cdf_u || !tgt_x_u ||
        B <= 0 || nt < 2 || nug < 1 || ntg < 1 || n_grids < 1 || tgt_rows < 1 ||
        (q != 0 && q != 2) || (pmask != WFOT_W1 && pmask != WFOT_W2 && !(pmask == WFOT_W12 && grad == nullptr)) || !(lambda > 0.0) ||
        (in_dtype != WFOT_F32 && in_dtype != WFOT_F64))
        return WFOT_ERR_INVALID_ARG;
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.t = t; a.w = w; a.dtype = in_dtype; a.t_stride = t_stride; a.nt = nt; a.grids = grids;
    a.n_grids = n_grids; a.B = B; a.nug = nug; a.ntg = ntg; a.lambda = lambda; a.q = q; a.pmask = pmask;
    a.transform = transform; a.tgt_cdf_t = tgt_cdf_t; a.tgt_x_t = tgt_x_t; a.tgt_cdf_u = tgt_cdf_u;
    a.tgt_x_u = tgt_x_u; a.tgt_rows = tgt_rows; a.W = W; a.grad = grad; a.dwg = dwg;
    a.status = status;
    return run_fused(a, workspace, workspace_bytes, (cudaStream_t)stream_);
}

int wfot_marginal_cdfs_batch(const void* t, const void* w, int in_dtype, long long t_stride, int nt,
                             const wfot_grid* grids, int n_grids, int B, int nug, int ntg, double lambda,
                             int q, int transform, double* cdf_t, double* cdf_u, double* amp,
                             void* workspace, size_t workspace_bytes, int32_t* status, void* stream_) {
    if (!t || !w || !grids || !cdf_t || !cdf_u || !workspace || B <= 0 || nt < 2 || nug < 1 || ntg < 1 ||
        n_grids < 1 || (q != 0 && q != 2) || !(lambda > 0.0) || (in_dtype != WFOT_F32 && in_dtype != WFOT_F64))
        return WFOT_ERR_INVALID_ARG;
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.t = t; a.w = w; a.dtype = in_dtype; a.t_stride = t_stride; a.nt = nt; a.grids = grids;
    a.n_grids = n_grids; a.B = B; a.nug = nug; a.ntg = ntg; a.lambda = lambda; a.q = q; a.pmask = WFOT_W2;
    a.transform = transform; a.tgt_rows = 1;
    a.out_cdf_t = cdf_t; a.out_cdf_u = cdf_u; a.out_amp = amp;
    a.status = status;
    return run_fused(a, workspace, workspace_bytes, (cudaStream_t)stream_);
}

}  // extern "C"
