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

#include <mutex>
#include <vector>

#include "wfot_fused.cuh"
#include "wfot_dev_options.h"

namespace wfot {

// ------------------------------------------------------------------ k_scan
template <int R, int T, int MINB, int NT = 256>
__global__ void __launch_bounds__(NT, MINB) k_scan(FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WFOT_SMEM_POINTERS(a.L);
    (void)s_margt; (void)s_margu; (void)s_Rt; (void)s_Ru; (void)s_xt; (void)s_xu; (void)s_cf; (void)s_E;
    (void)s_tk; (void)s_dx; (void)s_gbins; (void)s_posf; (void)s_queue; (void)s_colpart; (void)s_etab;
    const int tid = threadIdx.x, lane = tid & 31;
    const int npix = a.nug * a.ntg, S = a.nt - 1;
    SegTable tb{s_A, s_H, s_bbox, S, a.Spad, T};
    int degen = 0, tiles = 0;
    const bool vec = (a.ntg & 1) == 0;      // both pixels of a column pair exist and their 4 bytes are aligned
    for (int i = blockIdx.x; i < a.B;) {
        const int b = a.b0 + i;
        const wfot_grid g = a.grids[b % a.n_grids];
        if (tid == 0) { s_hdr->degenerate = 0; s_qcount[1] = 0; }
        __syncthreads();
        // this CTA's next window: asked for now, needed after the scan (the round trip to L2 hides behind it)
        if (tid == 0) s_qcount[2] = (int)gridDim.x + atomicAdd(a.next_window, 1);
        PrepOut po{s_pn, s_A, s_H, s_bbox, tb.tile, s_pxs, s_pys, s_hdr};
        prep_window(a.t, a.w, a.dtype, (long long)b * a.t_stride, (long long)b * a.nt, a.nt, g,
                    a.nug, a.ntg, a.transform, po, s_red, nullptr);
        __syncthreads();
        degen += (tid == 0) ? s_hdr->degenerate : 0;
        uint16_t* const out = a.scan_out + (size_t)i * npix;
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
                uint16_t* const dst = out + (size_t)iu * a.ntg + it0;
                if (vec) {
                    *reinterpret_cast<uint32_t*>(dst) = code[2 * r] | (code[2 * r + 1] << 16);
                } else {
                    dst[0] = (uint16_t)code[2 * r];
                    if (it0 + 1 < a.ntg) dst[1] = (uint16_t)code[2 * r + 1];
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
// 256 threads.  A warp owns the rows (warp, warp + 8, ...) of the window and walks each row 32 pixels at a time:
// every warp sees every column, so the warps of a CTA finish together whatever the waveform looks like (with a
// fixed column block per warp, the warps over the busy part of the waveform kept the others waiting at the
// barrier: 13 % of all warp time), and the marginals are accumulated on the way - the lane-strided row sum in a
// register, the row-group column sums (kRowGroups = 8 = warps) in shared memory - so the density itself is never
// stored or re-read.  A pixel whose near-ties span distant tiles (a handful per window) is resolved in place by
// the whole warp.
template <int MINB, int T, int NT = 256>
__global__ void __launch_bounds__(NT, MINB) k_resolve(FusedArgs a) {
    constexpr int NW = NT / 32;
    static_assert(kRowGroups % NW == 0, "a row group (row mod kRowGroups) belongs to one warp");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WFOT_SMEM_POINTERS(a.L);
    (void)s_bbox; (void)s_keys; (void)s_Rt; (void)s_Ru; (void)s_cf; (void)s_E; (void)s_queue;
    (void)s_tk; (void)s_dx; (void)s_gbins; (void)s_posf; (void)s_margt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npix = a.nug * a.ntg, S = a.nt - 1;
    const size_t slab = (size_t)blockIdx.x * npix;
    SegTable tb{s_A, s_H, s_bbox, S, a.Spad, T};
    int zero_dist = 0, slow = 0, common = 0;
    int* const counter = a.next_window + 16;
    load_exp_table(s_etab);      // visible behind the first barrier of the window loop
    long long ph[5] = {0, 0, 0, 0, 0};
    const bool timed = a.dbg_phase != nullptr && tid == 0;
    for (int i = blockIdx.x; i < a.B;) {
        const int b = a.b0 + i;
        const wfot_grid g = a.grids[b % a.n_grids];
        const long long tk0 = timed ? clock64() : 0;
        if (tid == 0) s_hdr->degenerate = 0;
        for (int c = tid; c < kRowGroups * a.ntg_pad; c += NT) s_colpart[c] = 0.0;
        __syncthreads();
        // this CTA's next window: asked for now, needed after the tail (the round trip to L2 hides behind P1)
        if (tid == 0) s_qcount[2] = (int)gridDim.x + atomicAdd(counter, 1);
        PrepOut po{s_pn, s_A, s_H, nullptr, tb.tile, s_pxs, s_pys, s_hdr};
        prep_window(a.t, a.w, a.dtype, (long long)b * a.t_stride, (long long)b * a.nt, a.nt, g,
                    a.nug, a.ntg, a.transform, po, s_red, nullptr);
        __syncthreads();
        const WinHdr hdr = *s_hdr;
        for (int c = tid; c < a.ntg; c += NT) s_xt[c] = lin_axis(hdr.T0, hdr.Tstep, hdr.Tlast, c, a.ntg);
        for (int c = tid; c < a.nug; c += NT) s_xu[c] = lin_axis(hdr.U0, hdr.Ustep, hdr.Ulast, c, a.nug);
        __syncthreads();

        // ---------------- P1 + P2 sums
        const long long tk1 = timed ? clock64() : 0;
        const uint16_t* const in = a.scan_out + (size_t)i * npix;
        int32_t* const dbg = a.dbg_iray ? a.dbg_iray + (size_t)b * npix : nullptr;
        unsigned nxt = (warp < a.nug && lane < a.ntg) ? __ldcs(in + (size_t)warp * a.ntg + lane) : 0u;
#pragma unroll 1
        for (int iu = warp; iu < a.nug; iu += NW) {
            const uint16_t* const inrow = in + (size_t)iu * a.ntg;
            double* const colp = s_colpart + (iu & (kRowGroups - 1)) * a.ntg_pad;
            const double pyd = s_xu[iu];
            const float pyl = s_pys[iu];
            double rowacc = 0.0;
#pragma unroll 1
            for (int c0 = 0; c0 < a.ntg; c0 += 32) {
                const int c = c0 + lane;
                const int it = min(c, a.ntg - 1);                     // lanes past the row end shadow its last pixel
                const bool live = c < a.ntg;
                const unsigned v = nxt;
                {   // next chunk of this row, or the first chunk of the warp's next row (its latency would otherwise
                    // be exposed once per row)
                    const bool last = c0 + 32 >= a.ntg;
                    const int cn = last ? lane : c + 32;
                    const uint16_t* const rn = last ? inrow + (size_t)NW * a.ntg : inrow;
                    if (cn < a.ntg && (!last || iu + NW < a.nug)) nxt = __ldcs(rn + cn);
                }
                const float pxl = s_pxs[it];
                const double pxd = s_xt[it];
                PixelHit hit;
                float kb1;
                const bool done = resolve_pixel_coded<T>(tb, s_pn, pxl, pyl, pxd, pyd, (int)(v & kScanTileMask),
                                                         (v & kScanFlag2) != 0u, (v & kScanFlag3) != 0u, hit, kb1);
                unsigned amb = __ballot_sync(0xffffffffu, live && !done);
                while (amb) {                                         // rare: all-segment rescan by the whole warp
                    const int src = __ffs((int)amb) - 1;
                    amb &= amb - 1u;
                    PixelHit h2;
                    resolve_pixel_warp(tb, s_pn, __shfl_sync(0xffffffffu, pxl, src), pyl,
                                       __shfl_sync(0xffffffffu, pxd, src), pyd, __shfl_sync(0xffffffffu, kb1, src), h2);
                    if (lane == src) { hit = h2; ++slow; }
                }
                const double pdf = store_pixel<false>(a, s_pn, s_etab, slab, it, iu, hit, pyd, zero_dist, dbg, live);
                if (live) {
                    rowacc += pdf;                                    // columns lane, lane + 32, ... in ascending order
                    colp[it] += pdf;                                  // rows of one group (row mod 8) in ascending order
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) rowacc += __shfl_xor_sync(0xffffffffu, rowacc, off);
            if (lane == 0) s_margu[iu] = rowacc;
        }
        __syncthreads();   // sums and scratch slab complete (block-scope visibility of global writes)
        long long tk3 = 0;
        const long long tk2 = timed ? clock64() : 0;
        common += window_tail<NT, true, (MINB == 2 ? 8 : 4)>(a, smem_raw, b, slab, hdr, timed ? &tk3 : nullptr);
        i = s_qcount[2];
        __syncthreads();
        if (timed) {
            const long long tk4 = clock64();
            ph[0] += tk1 - tk0; ph[1] += tk2 - tk1; ph[2] += tk3 - tk2; ph[3] += tk4 - tk3; ph[4] += 1;
        }
    }
    if (a.status) {
        if (zero_dist) atomicAdd(a.status + WFOT_STAT_ZERO_DIST, zero_dist);
        if (slow) atomicAdd(a.status + WFOT_STAT_SLOW_PIXELS, slow);
        if (common) atomicAdd(a.status + WFOT_STAT_COMMON_CDF, common);
    }
    if (timed)
        for (int k = 0; k < 5; ++k) atomicAdd(a.dbg_phase + k, (unsigned long long)ph[k]);
}

// ------------------------------------------------------------------ host side
// The batch is cut into chunks; chunk c is scanned on the caller's stream and resolved on a helper stream.  Both
// kernels are launched with their full stand-alone grids, so they do not share SMs while both have work (measured:
// reduced grids sized for co-residency - one scan CTA + two resolve CTAs per SM - lose 10-15 %); what the second
// stream buys is that the scan of chunk c + 1 moves onto SMs as the resolve CTAs of chunk c retire, i.e. the
// tails of the persistent kernels (up to one window per CTA, ~7 % per chunk) are filled with the next chunk's work.
// Three chunk-sized scan-result buffers rotate; events order scan(c) -> resolve(c) -> scan(c + 3).  The helper
// stream joins the caller's stream before the call returns, so the call stays stream-ordered for the caller.
constexpr int kScanBuffers = 3;
constexpr int kMaxChunks = 30;           // two window counters per chunk in the 256-byte counter block

static bool overlap_enabled() { return dev_option(kOptOverlap) != 1; }

static int split_chunk(int B, int nug, int ntg, int sms) {
    int c;
    if (const int o = dev_option(kOptSplitChunk)) c = o;
    else {
        // up to 2.5 GiB of scan results per buffer (2 bytes per pixel), between 8 and 64 windows per SM.  Measured on cfg5 (9472 windows):
        // one launch pair 369 k evals/s, two to five overlapped pairs 363-365 k - the tails that the second stream
        // fills cost less than the kernels lose while they share the SMs - so chunks are as large as memory allows
        const long long by_bytes = (2560LL << 20) / ((long long)nug * ntg * 2);
        c = (int)(by_bytes < 8LL * sms ? 8LL * sms : by_bytes > 64LL * sms ? 64LL * sms : by_bytes);
    }
    if (c < (B + kMaxChunks - 1) / kMaxChunks) c = (B + kMaxChunks - 1) / kMaxChunks;
    if (c >= B) return B;
    const int n = (B + c - 1) / c;        // equal chunks: no short last one
    return (B + n - 1) / n;
}

bool split_wanted(int B, int nt, int nug, int ntg, int sms) {
    if (dev_option(kOptPipeline) == 1) return false;
    if (dev_option(kOptPipeline) == 2) return seg_pad(nt) / kTileMin < 16384;
    // at least four windows per SM (smaller batches: the single-kernel form, with thread-block clusters for the smallest)
    // and rows of at least one warp's width
    // (16-bit scan results hold tile indices below 16384)
    return ntg >= 32 && (long long)nug * ntg >= 2048 && B >= 4 * sms && seg_pad(nt) / kTileMin < 16384;
}

size_t split_workspace_bytes(int B, int nt, int nug, int ntg, int sms) {
    if (!split_wanted(B, nt, nug, ntg, sms)) return 0;
    const int chunk = split_chunk(B, nug, ntg, sms);
    const int nbuf = (B + chunk - 1) / chunk < kScanBuffers ? (B + chunk - 1) / chunk : kScanBuffers;
    return (size_t)nbuf * ((((size_t)chunk * nug * ntg * 2) + 255) & ~(size_t)255) + 512;
}

// Helper stream + events per (device, caller stream); created on first use, kept for the life of the process.
struct SideLane {
    int dev; cudaStream_t user; cudaStream_t side;
    cudaEvent_t scanned[kScanBuffers], resolved[kScanBuffers];
};
static std::mutex g_lane_mutex;
static std::vector<SideLane> g_lanes;

static SideLane* side_lane(cudaStream_t user) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(g_lane_mutex);
    for (auto& l : g_lanes)
        if (l.dev == dev && l.user == user) return &l;
    if (g_lanes.size() >= 64) return nullptr;        // callers with many streams: sequential form
    SideLane l;
    l.dev = dev; l.user = user;
    if (cudaStreamCreateWithFlags(&l.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (int i = 0; i < kScanBuffers; ++i)
        if (cudaEventCreateWithFlags(&l.scanned[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&l.resolved[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    g_lanes.reserve(64);
    g_lanes.push_back(l);
    return &g_lanes.back();
}

template <int T>
static int launch_split_t(FusedArgs a, unsigned char* ws, size_t ws_bytes, cudaStream_t stream) {
    const size_t npix = (size_t)a.nug * a.ntg;
    int sms = wfot_device_sm_count();
    if (sms <= 0) return cuda_fail(cudaGetLastError(), "device query");
    FusedArgs as = a, ar = a;
    as.L = make_layout(a.nt, a.Spad, a.ntg_pad, a.nug_pad, a.nmax, kLayoutScan);
    // resolve kernel shape: 128 registers x 2 CTAs per SM for long windows (deeper unrolled gradient assembly: its slab
    // read-back is the latency sink of the kernel), 80 registers x 3 per SM for short ones; the layout gets the second
    // OT scratch set (the two marginal problems on two warps) unless that costs a resident CTA
    int rshape = dev_option(kOptResolveShape);
    {
        const int base = make_layout(a.nt, a.Spad, a.ntg_pad, a.nug_pad, a.nmax, kLayoutResolve).total;
        if (rshape != 1 && rshape != 2 && rshape != 5) rshape = base > 48 * 1024 ? 1 : 2;
        ar.L = make_layout_auto(a.nt, a.Spad, a.ntg_pad, a.nug_pad, a.nmax, kLayoutResolve, rshape == 1 ? 2 : rshape == 2 ? 3 : 6);
    }
    const size_t smem_s = (size_t)as.L.total, smem_r = (size_t)ar.L.total;
    // overlap needs a helper stream; not while the caller's stream is being captured into a CUDA graph
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cap);
    SideLane* lane = (overlap_enabled() && cap == cudaStreamCaptureStatusNone) ? side_lane(stream) : nullptr;
    int per_sm = 0;
    const int scan3 = dev_option(kOptScanShape) != 2;             // default: 80 registers, 3 CTAs per SM (2: 128 x 2; 4: 128 threads x 6)
    // windows with fewer footprints than a 256-thread CTA has warps to spare (79 x 61 pixels: 20 footprints for 8 warps,
    // and the barriers of the window preparation in between): 128-thread CTAs, six per SM
    const int nfoot = max_footprints<4>(a.ntg, a.nug);
    const int sshape = dev_option(kOptScanShape);
    const bool scan128 = sshape == 4 || (sshape == 0 && nfoot <= 32);
    const int sthreads = scan128 ? 128 : 256;
    const int scan_ctas = scan128 ? resident_ctas(k_scan<4, T, 6, 128>, smem_s, &per_sm, 128)
                        : scan3 ? resident_ctas(k_scan<4, T, 3>, smem_s, &per_sm, 256)
                                : resident_ctas(k_scan<4, T, 2>, smem_s, &per_sm, 256);
    if (scan_ctas < 1) return cuda_fail(cudaGetLastError(), "k_scan occupancy");
    const int rthreads = rshape == 5 ? 128 : 256;
    const int res_ctas = rshape == 1 ? resident_ctas(k_resolve<2, T, 256>, smem_r, &per_sm, 256)
                       : rshape == 2 ? resident_ctas(k_resolve<3, T, 256>, smem_r, &per_sm, 256)
                                     : resident_ctas(k_resolve<6, T, 128>, smem_r, &per_sm, 128);
    if (res_ctas < 1) return cuda_fail(cudaGetLastError(), "k_resolve occupancy");
    const size_t slab_px = 16;                         // bytes per pixel of the per-CTA scratch slab
    const int Btot = a.B;
    const int chunk = split_chunk(Btot, a.nug, a.ntg, sms);
    const int nchunks = (Btot + chunk - 1) / chunk;
    const int nbuf = nchunks < kScanBuffers ? nchunks : kScanBuffers;
    // workspace: [scan results: nbuf chunks][slabs of the resolve CTAs]
    uintptr_t p = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
    const size_t scan_bytes = ((size_t)chunk * npix * 2 + 255) & ~(size_t)255;
    if (ws_bytes < (p - (uintptr_t)ws) + nbuf * scan_bytes + npix * slab_px) return WFOT_ERR_WORKSPACE;
    unsigned char* const scan_base = (unsigned char*)p;
    p += nbuf * scan_bytes;
    const size_t max_ctas = (ws_bytes - (p - (uintptr_t)ws)) / (npix * slab_px);
    const int res_grid = (size_t)res_ctas > max_ctas ? (int)max_ctas : res_ctas;
    unsigned char* q = (unsigned char*)p;
    ar.s_pdf = nullptr;      // the resolve kernel sums the density on the fly
    ar.s_w = (ulonglong2*)q;
    as.cluster = ar.cluster = 1;
    int* const counters = a.next_window;             // 64 ints, zeroed by the caller on `stream`
    cudaStream_t rstream = lane ? lane->side : stream;
    for (int c = 0; c < nchunks; ++c) {
        const int c0 = c * chunk;
        const int nb = (Btot - c0 < chunk) ? Btot - c0 : chunk;
        const int buf = c % nbuf;
        as.b0 = ar.b0 = c0; as.B = ar.B = nb;
        as.scan_out = ar.scan_out = (uint16_t*)(scan_base + (size_t)buf * scan_bytes);
        as.next_window = counters + 2 * c;
        ar.next_window = counters + 2 * c + 1 - 16;  // k_resolve counts at next_window + 16
        if (lane && c >= nbuf && cudaStreamWaitEvent(stream, lane->resolved[buf], 0) != cudaSuccess)
            return cuda_fail(cudaGetLastError(), "cudaStreamWaitEvent");
        const int skip = dev_option(kOptSkipKernel);
        if (skip == 2) {}
        else if (scan128) k_scan<4, T, 6, 128><<<scan_ctas < nb ? scan_ctas : nb, sthreads, smem_s, stream>>>(as);
        else if (scan3) k_scan<4, T, 3><<<scan_ctas < nb ? scan_ctas : nb, 256, smem_s, stream>>>(as);
        else k_scan<4, T, 2><<<scan_ctas < nb ? scan_ctas : nb, 256, smem_s, stream>>>(as);
        if (lane) {
            if (cudaEventRecord(lane->scanned[buf], stream) != cudaSuccess ||
                cudaStreamWaitEvent(rstream, lane->scanned[buf], 0) != cudaSuccess)
                return cuda_fail(cudaGetLastError(), "cudaEventRecord");
        }
        const int rc = res_grid < nb ? res_grid : nb;
        if (skip == 1) {}
        else if (rshape == 1) k_resolve<2, T, 256><<<rc, 256, smem_r, rstream>>>(ar);
        else if (rshape == 2) k_resolve<3, T, 256><<<rc, 256, smem_r, rstream>>>(ar);
        else k_resolve<6, T, 128><<<rc, rthreads, smem_r, rstream>>>(ar);
        note_launches(2);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "wfot_misfit_grad_batch (scan + resolve) launch");
        if (lane && cudaEventRecord(lane->resolved[buf], rstream) != cudaSuccess)
            return cuda_fail(cudaGetLastError(), "cudaEventRecord");
    }
    if (lane) {                                      // join: later work on the caller's stream sees every result
        for (int i = 0; i < nbuf; ++i)
            if (cudaStreamWaitEvent(stream, lane->resolved[i], 0) != cudaSuccess)
                return cuda_fail(cudaGetLastError(), "cudaStreamWaitEvent");
    }
    return WFOT_OK;
}

int launch_split(FusedArgs a, unsigned char* ws, size_t ws_bytes, cudaStream_t stream) {
    return tile_for(a.nt) == 8 ? launch_split_t<8>(a, ws, ws_bytes, stream)
                               : launch_split_t<16>(a, ws, ws_bytes, stream);
}

}  // namespace wfot
