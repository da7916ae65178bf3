// wfot_split.cu -- the throughput path as TWO kernels, for large batches of large windows.
//
// The single-kernel form (wfot_fused.cu) runs the FP32 segment scan (ILP-rich, 120+ registers) and the
// FP64 per-pixel work (serial dependency chains: exact candidate evaluation, sqrt/exp/divide of the
// density, gradient weights) in the same threads, so the FP64 half is pinned at the scan's occupancy
// (16 warps per SM) and spends most of its cycles waiting on its own previous instruction.  Here:
//
//   k_scan<R, T>     P0 prep_window + the pruned best-first scan of every warp footprint.  Per pixel it
//                    leaves 8 bytes in a scratch array: the FP32 minimum b1 and the tile that holds it,
//                    plus two flags (another / several other tiles within the FP32 rounding tolerance).
//   k_resolve<NT,..> P0 prep_window again (bit-identical tables, 1 % of the work) + one pixel per thread
//                    in row-major order: FP32 re-evaluation of the winning tile, FP64 reference-order
//                    evaluation of the candidates, density and gradient weights to the per-CTA slab,
//                    then window_tail() (marginals, OT, gradient assembly) as in the single-kernel form.
//                    Runs at 64-85 registers, 24-32 warps per SM; neighbouring lanes hold neighbouring
//                    pixels, so their tile loads coalesce into shared-memory broadcasts and their slab
//                    stores into full 128-byte lines.
//
// Both kernels are persistent (windows drawn from a global counter).  Results are bit-identical to the
// single-kernel form: same prep, same scan, same candidate set, same FP64 evaluation order.
#include <cuda_runtime.h>
#include <string.h>

#include "wfot_fused.cuh"
#include "wfot_dev_options.h"

namespace wfot {

// ------------------------------------------------------------------ k_scan
template <int R, int T>
__global__ void __launch_bounds__(256, 2) k_scan(FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WFOT_SMEM_POINTERS(a.L);
    (void)s_margt; (void)s_margu; (void)s_Rt; (void)s_Ru; (void)s_xt; (void)s_xu; (void)s_cf; (void)s_E;
    (void)s_tk; (void)s_dx; (void)s_gbins; (void)s_posf; (void)s_queue;
    const int tid = threadIdx.x, lane = tid & 31;
    const int npix = a.nug * a.ntg, S = a.nt - 1;
    SegTable tb{s_A, s_H, s_bbox, S, a.Spad, T};
    int degen = 0, tiles = 0;
    const bool vec = (a.ntg & 1) == 0;      // both pixels of a column pair exist and their 16 bytes are aligned
    for (int i = blockIdx.x; i < a.B;) {
        const int b = a.b0 + i;
        const wfot_grid g = a.grids[b % a.n_grids];
        if (tid == 0) {
            s_hdr->degenerate = 0; s_qcount[1] = 0;
            s_qcount[2] = (int)gridDim.x + atomicAdd(a.next_window, 1);          // this CTA's next window
        }
        __syncthreads();
        PrepOut po{s_pn, s_A, s_H, s_bbox, tb.tile, s_pxs, s_pys, s_hdr};
        prep_window(a.t, a.w, a.dtype, (long long)b * a.t_stride, (long long)b * a.nt, a.nt, g,
                    a.nug, a.ntg, a.transform, po, s_red, nullptr);
        __syncthreads();
        degen += (tid == 0) ? s_hdr->degenerate : 0;
        uint2* const out = a.scan_out + (size_t)i * npix;
        const FootMap fm = make_footmap<R>(a.ntg, a.nug, fabsf(s_pxs[a.ntg - 1] - s_pxs[0]),
                                           fabsf(s_pys[a.nug - 1] - s_pys[0]));
        for (;;) {
            int f = 0;
            if (lane == 0) f = atomicAdd(s_qcount + 1, 1);
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
            unsigned code[2 * R];
#pragma unroll
            for (int k = 0; k < 2 * R; ++k) {
                const float thr = b1[k] + tau32(b1[k]);
                code[k] = (unsigned)t1[k] | (!(b2[k] > thr) ? kScanFlag2 : 0u) | ((b3[k] <= thr) ? kScanFlag3 : 0u);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int iu = rg * R + r;
                if (iu >= a.nug) break;
                uint2* const dst = out + (size_t)iu * a.ntg + it0;
                if (vec) {
                    *reinterpret_cast<uint4*>(dst) = make_uint4(__float_as_uint(b1[2 * r]), code[2 * r],
                                                                __float_as_uint(b1[2 * r + 1]), code[2 * r + 1]);
                } else {
                    dst[0] = make_uint2(__float_as_uint(b1[2 * r]), code[2 * r]);
                    if (it0 + 1 < a.ntg) dst[1] = make_uint2(__float_as_uint(b1[2 * r + 1]), code[2 * r + 1]);
                }
            }
        }
        i = s_qcount[2];
        __syncthreads();     // every warp is done with this window's tables before the next prep overwrites them
    }
    if (a.status) {
        if (degen) atomicAdd(a.status + WFOT_STAT_DEGENERATE_SEG, degen);
        if (lane == 0 && tiles)
            atomicAdd(reinterpret_cast<unsigned long long*>(a.status + WFOT_STAT_SCAN_TILES),
                      (unsigned long long)tiles * (R / 4) * (T / 8));
    }
}

// ------------------------------------------------------------------ k_resolve
template <int NT, int MINB, int T>
__global__ void __launch_bounds__(NT, MINB) k_resolve(FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WFOT_SMEM_POINTERS(a.L);
    (void)s_bbox; (void)s_keys; (void)s_margt; (void)s_margu; (void)s_Rt; (void)s_Ru; (void)s_cf; (void)s_E;
    (void)s_tk; (void)s_dx; (void)s_gbins; (void)s_posf;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npix = a.nug * a.ntg, S = a.nt - 1;
    const size_t slab = (size_t)blockIdx.x * npix;
    SegTable tb{s_A, s_H, s_bbox, S, a.Spad, T};
    int zero_dist = 0, slow = 0, common = 0;
    int* const counter = a.next_window + 16;
    const int dit = NT % a.ntg, diu = NT / a.ntg;
    for (int i = blockIdx.x; i < a.B;) {
        const int b = a.b0 + i;
        const wfot_grid g = a.grids[b % a.n_grids];
        if (tid == 0) {
            s_hdr->degenerate = 0; s_qcount[0] = 0;
            s_qcount[2] = (int)gridDim.x + atomicAdd(counter, 1);
        }
        if (a.grad) {      // P4 accumulates into these rows with L2 reductions
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
        for (int c = tid; c < a.ntg; c += NT) s_xt[c] = lin_axis(hdr.T0, hdr.Tstep, hdr.Tlast, c, a.ntg);
        for (int c = tid; c < a.nug; c += NT) s_xu[c] = lin_axis(hdr.U0, hdr.Ustep, hdr.Ulast, c, a.nug);
        __syncthreads();

        // ---------------- P1: one pixel per thread, row-major
        const uint2* const in = a.scan_out + (size_t)i * npix;
        int it = tid % a.ntg, iu = tid / a.ntg;
        uint2 nxt = (tid < npix) ? __ldcs(in + tid) : make_uint2(0u, 0u);
#pragma unroll 1
        for (int pix = tid; pix < npix; pix += NT) {
            const uint2 v = nxt;
            if (pix + NT < npix) nxt = __ldcs(in + pix + NT);
            const float kb1 = __uint_as_float(v.x);
            const float thr = kb1 + tau32(kb1);
            const double pyd = s_xu[iu];
            PixelHit hit;
            bool done = resolve_pixel_flagged<T>(tb, s_pn, s_pxs[it], s_pys[iu], s_xt[it], pyd, thr,
                                                 (int)(v.y & kScanTileMask), (v.y & kScanFlag2) != 0u,
                                                 (v.y & kScanFlag3) != 0u, hit);
            if (!done) {
                const int qi = atomicAdd(s_qcount, 1);
                if (qi < kFQCap) {
                    s_queue[qi] = FQEntry{pix, kb1};
                } else {
                    ++slow;
                    resolve_pixel_full(tb, s_pn, s_pxs[it], s_pys[iu], s_xt[it], pyd, kb1, hit);
                    done = true;
                }
            }
            if (done) store_pixel(a, s_pn, slab, it, iu, hit, pyd, zero_dist);
            it += dit; iu += diu;
            if (it >= a.ntg) { it -= a.ntg; ++iu; }
        }
        __syncthreads();
        {
            const int nq = min(*s_qcount, kFQCap);
            for (int e = warp; e < nq; e += NT / 32) {
                const FQEntry qe = s_queue[e];
                const int qt = qe.pix % a.ntg, qu = qe.pix / a.ntg;
                PixelHit hit;
                resolve_pixel_warp(tb, s_pn, s_pxs[qt], s_pys[qu], s_xt[qt], s_xu[qu], qe.b1, hit);
                if (lane == 0) { store_pixel(a, s_pn, slab, qt, qu, hit, s_xu[qu], zero_dist); ++slow; }
            }
        }
        __syncthreads();   // scratch slab complete (block-scope visibility of global writes)
        common += window_tail<NT>(a, smem_raw, b, slab, hdr);
        i = s_qcount[2];
        __syncthreads();
    }
    if (a.status) {
        if (zero_dist) atomicAdd(a.status + WFOT_STAT_ZERO_DIST, zero_dist);
        if (slow) atomicAdd(a.status + WFOT_STAT_SLOW_PIXELS, slow);
        if (common) atomicAdd(a.status + WFOT_STAT_COMMON_CDF, common);
    }
}

// ------------------------------------------------------------------ host side
// Windows per scan/resolve launch pair: the scratch array holds 8 bytes per pixel of one chunk.
static int split_chunk(int B, int nug, int ntg) {
    if (const int o = dev_option(kOptSplitChunk)) return o < B ? o : B;
    const long long cap = (4LL << 30) / ((long long)nug * ntg * 8);        // <= 4 GiB of scan results in flight
    long long c = cap < 1 ? 1 : cap;
    return (int)(c < B ? c : B);
}

bool split_wanted(int B, int nt, int nug, int ntg, int sms) {
    (void)nt;
    if (dev_option(kOptPipeline) == 1) return false;
    if (dev_option(kOptPipeline) == 2) return true;
    // large windows (the single-kernel form would run 256-thread CTAs) and at least four windows per SM
    return false && (long long)nug * ntg > 16384 && B >= 4 * sms;
}

size_t split_workspace_bytes(int B, int nt, int nug, int ntg, int sms) {
    if (!split_wanted(B, nt, nug, ntg, sms)) return 0;
    return (size_t)split_chunk(B, nug, ntg) * nug * ntg * 8 + 256;
}

template <int T>
static int launch_split_t(FusedArgs a, unsigned char* ws, size_t ws_bytes, cudaStream_t stream) {
    const size_t smem = (size_t)a.L.total;
    const size_t npix = (size_t)a.nug * a.ntg;
    int per_sm = 0;
    const int scan_ctas = resident_ctas(k_scan<4, T>, smem, &per_sm, 256);
    if (scan_ctas < 1) return cuda_fail(cudaGetLastError(), "k_scan occupancy");
    int shape = dev_option(kOptResolveShape);
    if (shape < 1 || shape > 3) shape = 2;
    int res_ctas = -1, res_threads = 256;
    if (shape == 1) res_ctas = resident_ctas(k_resolve<256, 2, T>, smem, &per_sm, 256);
    else if (shape == 2) res_ctas = resident_ctas(k_resolve<256, 3, T>, smem, &per_sm, 256);
    else { res_threads = 512; res_ctas = resident_ctas(k_resolve<512, 2, T>, smem, &per_sm, 512); }
    if (res_ctas < 1) return cuda_fail(cudaGetLastError(), "k_resolve occupancy");
    const int Btot = a.B;
    const int chunk = split_chunk(Btot, a.nug, a.ntg);
    // workspace: [scan results of one chunk][slabs of the resolve CTAs]
    uintptr_t p = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
    const size_t scan_bytes = (size_t)chunk * npix * 8;
    if (ws_bytes < (p - (uintptr_t)ws) + scan_bytes + npix * 28) return WFOT_ERR_WORKSPACE;
    a.scan_out = (uint2*)p;
    p += scan_bytes;
    const size_t max_ctas = (ws_bytes - (p - (uintptr_t)ws)) / (npix * 28);
    if ((size_t)res_ctas > max_ctas) res_ctas = (int)max_ctas;
    unsigned char* q = (unsigned char*)p;
    a.s_pdf = (double*)q;   q += (size_t)res_ctas * npix * 8;
    a.s_wa = (double*)q;    q += (size_t)res_ctas * npix * 8;
    a.s_wb = (double*)q;    q += (size_t)res_ctas * npix * 8;
    a.s_idx = (int32_t*)q;
    a.cluster = 1;
    for (int c0 = 0; c0 < Btot; c0 += chunk) {
        const int nb = (Btot - c0 < chunk) ? Btot - c0 : chunk;
        a.b0 = c0; a.B = nb;
        if (c0 > 0 && cudaMemsetAsync(a.next_window, 0, 256, stream) != cudaSuccess)
            return cuda_fail(cudaGetLastError(), "cudaMemsetAsync");
        k_scan<4, T><<<scan_ctas < nb ? scan_ctas : nb, 256, smem, stream>>>(a);
        const int rc = res_ctas < nb ? res_ctas : nb;
        if (shape == 1) k_resolve<256, 2, T><<<rc, 256, smem, stream>>>(a);
        else if (shape == 2) k_resolve<256, 3, T><<<rc, 256, smem, stream>>>(a);
        else k_resolve<512, 2, T><<<rc, res_threads, smem, stream>>>(a);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "wfot_misfit_grad_batch (scan + resolve) launch");
    }
    return WFOT_OK;
}

int launch_split(FusedArgs a, unsigned char* ws, size_t ws_bytes, cudaStream_t stream) {
    return tile_for(a.nt) == 8 ? launch_split_t<8>(a, ws, ws_bytes, stream)
                               : launch_split_t<16>(a, ws, ws_bytes, stream);
}

}  // namespace wfot
