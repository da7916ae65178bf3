// wfot_fused.cu -- the throughput path: one persistent kernel that takes a
// waveform window from raw samples to (W^t, W^u, dW^t/dw, dW^u/dw, dW^t/dx0)
// without the 2-D field ever leaving the chip's caches as an HBM-sized array.
//
// One CTA owns one window at a time (grid = resident CTAs, windows strided):
//   P0  prep_window            FP64 normalisation + FP32 segment table -> shared memory
//   P1  scan_block + resolve   nearest segment per pixel (FP32 brute force, FP64 exact
//                              tie resolution); per pixel {iray, pdf, wa, wb} go to a
//                              per-CTA scratch slab (28 B/pixel, L2-resident, re-used
//                              for every window the CTA processes)
//   P2  marginals              fixed-order column / row sums of pdf -> time / amplitude
//                              marginals of the normalised density (OTlib.py:92-93,155-156)
//   P3  block_ot1d x 2         CDF scan, merge, W_p^p, dW/df, dW/dx0 per marginal
//                              (OTlib.py:596-706) and <dW, pbar> (OTlib.py:1141,1144-1145)
//   P4  gradient assembly      sum_k pdf_k (R_k - Rbar)/A dd_k/dw_j keyed by iray
//                              (FingerprintLib.py:205-228), run-combined per pixel column
// Reference chain replaced: ricker_util.py:386-388 (BuildOTobjfromWaveform ->
// CalcWasserWaveform(deriv=True, returnmarg=True)).
#include <cuda_runtime.h>
#include <stdio.h>

#include "wfot_device.cuh"
#include "wfot_host.h"
#include "wfot_ot.cuh"

namespace wfot {

constexpr int kFQCap = 1024;
struct FQEntry { int pix; float b1; };

struct FusedArgs {
    const void* t; const void* w; int dtype; long long t_stride; int nt;
    const wfot_grid* grids; int n_grids; int B; int nug, ntg;
    double lambda; int q, pmask, transform;
    const double* tgt_cdf_t; const double* tgt_x_t; const double* tgt_cdf_u; const double* tgt_x_u;
    int tgt_per_window;
    double* W; double* grad; double* dwg;
    // per-CTA scratch slabs
    double* s_pdf; double* s_wa; double* s_wb; int32_t* s_idx;
    int32_t* status;
    int Spad, ntg_pad, nug_pad, nmax;
};

struct FusedSmem {
    double2* pn; float4* A; float4* B; float* H; float* pxs; float* pys;
    double* margt; double* margu; double* Rt; double* Ru; double* xt; double* xu;
    double* cf; double* tk; double* dx; double* E; double* red; double* gbins;
    int* posf; FQEntry* queue; WinHdr* hdr; int* qcount;
};

__host__ __device__ inline size_t fused_smem_bytes(int nt, int Spad, int ntg_pad, int nug_pad, int nmax) {
    size_t s = 0;
    s += (size_t)nt * 16;                 // pn
    s += (size_t)Spad * 36;               // A, B, H
    s += (size_t)(ntg_pad + nug_pad) * 4; // pxs, pys
    s += (size_t)(ntg_pad + nug_pad) * 8 * 3;   // marg, R, x
    s += (size_t)nmax * 8 * 2 + (size_t)(2 * nmax) * 8 * 2;   // cf, E, tk, dx
    s += 64 * 8;                          // red
    s += (size_t)2 * nt * 8;              // gbins
    s += (size_t)nmax * 4;                // posf
    s += (size_t)kFQCap * sizeof(FQEntry);
    s += 128 + 16;                        // hdr, qcount
    return s + 64;
}

__device__ __forceinline__ FusedSmem carve(unsigned char* p, const FusedArgs& a) {
    FusedSmem s;
    s.pn = (double2*)p;      p += (size_t)a.nt * 16;
    s.A = (float4*)p;        p += (size_t)a.Spad * 16;
    s.B = (float4*)p;        p += (size_t)a.Spad * 16;
    s.margt = (double*)p;    p += (size_t)a.ntg_pad * 8;
    s.margu = (double*)p;    p += (size_t)a.nug_pad * 8;
    s.Rt = (double*)p;       p += (size_t)a.ntg_pad * 8;
    s.Ru = (double*)p;       p += (size_t)a.nug_pad * 8;
    s.xt = (double*)p;       p += (size_t)a.ntg_pad * 8;
    s.xu = (double*)p;       p += (size_t)a.nug_pad * 8;
    s.cf = (double*)p;       p += (size_t)a.nmax * 8;
    s.E = (double*)p;        p += (size_t)a.nmax * 8;
    s.tk = (double*)p;       p += (size_t)a.nmax * 16;
    s.dx = (double*)p;       p += (size_t)a.nmax * 16;
    s.red = (double*)p;      p += 64 * 8;
    s.gbins = (double*)p;    p += (size_t)a.nt * 16;
    s.hdr = (WinHdr*)p;      p += 128;
    s.queue = (FQEntry*)p;   p += (size_t)kFQCap * sizeof(FQEntry);
    s.H = (float*)p;         p += (size_t)a.Spad * 4;
    s.pxs = (float*)p;       p += (size_t)a.ntg_pad * 4;
    s.pys = (float*)p;       p += (size_t)a.nug_pad * 4;
    s.posf = (int*)p;        p += (size_t)a.nmax * 4;
    s.qcount = (int*)p;
    return s;
}

// pixel -> scratch: density and the two gradient weights of its nearest segment
__device__ __forceinline__ void store_pixel(const FusedArgs& a, const FusedSmem& sm, size_t slab,
                                            int it, int iu, const PixelHit& hit, double py, int& zero_dist) {
    const PixelVals v = pixel_values(sm.pn, hit, py, a.lambda, a.q);
    const size_t k = slab + (size_t)iu * a.ntg + it;
    double wgt = v.pdf * v.g;                                  // pdf * dddx_y (FingerprintLib.py:355)
    if (a.q == 2) wgt *= 2.0 * fabs(v.d);                      // :214-217
    zero_dist += (v.d == 0.0);
    a.s_pdf[k] = v.pdf;
    a.s_wa[k] = (1.0 - hit.lam) * wgt;                         // -> sample iray    (:223)
    a.s_wb[k] = hit.lam * wgt;                                 // -> sample iray+1  (:224)
    a.s_idx[k] = hit.s;
}

template <int R>
__global__ void __launch_bounds__(256) k_misfit_grad(FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FusedSmem sm = carve(smem_raw, a);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npix = a.nug * a.ntg, S = a.nt - 1;
    const size_t slab = (size_t)blockIdx.x * npix;
    const int ncp = (a.ntg + 1) >> 1, nrg = (a.nug + R - 1) / R, nblk = ncp * nrg;
    const SegTable tb{sm.A, sm.B, sm.H, S, a.Spad};
    int zero_dist = 0, slow = 0, common = 0, degen = 0;

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        // ---------------- P0: window -> shared memory
        const wfot_grid g = a.grids[a.n_grids == 1 ? 0 : b];
        if (tid == 0) { sm.hdr->degenerate = 0; *sm.qcount = 0; }
        __syncthreads();
        PrepOut po{sm.pn, sm.A, sm.B, sm.H, sm.pxs, sm.pys, sm.hdr};
        prep_window(a.t, a.w, a.dtype, (long long)b * a.t_stride, (long long)b * a.nt, a.nt, g,
                    a.nug, a.ntg, a.transform, po, sm.red, nullptr);
        __syncthreads();
        const WinHdr hdr = *sm.hdr;
        degen += (tid == 0) ? hdr.degenerate : 0;
        for (int i = tid; i < a.ntg; i += 256) sm.xt[i] = lin_axis(hdr.T0, hdr.Tstep, hdr.Tlast, i, a.ntg);
        for (int i = tid; i < a.nug; i += 256) sm.xu[i] = lin_axis(hdr.U0, hdr.Ustep, hdr.Ulast, i, a.nug);
        for (int j = tid; j < 2 * a.nt; j += 256) sm.gbins[j] = 0.0;

        // ---------------- P1: nearest segment per pixel
        for (int blk = tid; blk < nblk; blk += 256) {
            const int cp = blk % ncp, rg = blk / ncp;
            const int it0 = 2 * cp, it1 = min(2 * cp + 1, a.ntg - 1);
            float py[R];
#pragma unroll
            for (int r = 0; r < R; ++r) py[r] = sm.pys[min(rg * R + r, a.nug - 1)];
            float b1[2 * R], b2[2 * R];
            int t1[2 * R];
            scan_block<R>(tb, sm.pxs[it0], sm.pxs[it1], py, b1, t1, b2);
            float lb1[2 * R], lb2[2 * R];
            int lt1[2 * R];
#pragma unroll
            for (int k = 0; k < 2 * R; ++k) { lb1[k] = b1[k]; lb2[k] = b2[k]; lt1[k] = t1[k]; }
#pragma unroll 1
            for (int k = 0; k < 2 * R; ++k) {
                const int it = 2 * cp + (k & 1), iu = rg * R + (k >> 1);
                if (it >= a.ntg || iu >= a.nug) continue;
                const float kb1 = lb1[k];
                const double pyd = sm.xu[iu];
                PixelHit hit;
                if (lb2[k] <= kb1 + tau32(kb1)) {
                    const int qi = atomicAdd(sm.qcount, 1);
                    if (qi < kFQCap) { sm.queue[qi] = FQEntry{iu * a.ntg + it, kb1}; continue; }
                    ++slow;
                    resolve_pixel_full(tb, sm.pn, sm.pxs[it], sm.pys[iu], sm.xt[it], pyd, kb1, hit);
                } else {
                    resolve_pixel(tb, sm.pn, sm.pxs[it], sm.pys[iu], sm.xt[it], pyd, kb1, lt1[k], hit);
                }
                store_pixel(a, sm, slab, it, iu, hit, pyd, zero_dist);
            }
        }
        __syncthreads();
        {
            const int nq = min(*sm.qcount, kFQCap);
            for (int e = warp; e < nq; e += 8) {
                const FQEntry qe = sm.queue[e];
                const int it = qe.pix % a.ntg, iu = qe.pix / a.ntg;
                PixelHit hit;
                resolve_pixel_warp(tb, sm.pn, sm.pxs[it], sm.pys[iu], sm.xt[it], sm.xu[iu], qe.b1, hit);
                if (lane == 0) { store_pixel(a, sm, slab, it, iu, hit, sm.xu[iu], zero_dist); ++slow; }
            }
        }
        __syncthreads();   // scratch slab complete (block-scope visibility of global writes)

        // ---------------- P2: marginals of the normalised density
        double part = 0.0;
        for (int c = tid; c < a.ntg; c += 256) {
            double s0 = 0.0;
            for (int iu = 0; iu < a.nug; ++iu) s0 += a.s_pdf[slab + (size_t)iu * a.ntg + c];
            sm.margt[c] = s0;
            part += s0;
        }
        const double A = block_sum(part, sm.red);                       // OTpdf.amp (OTlib.py:92)
        for (int iu = warp; iu < a.nug; iu += 8) {
            double s0 = 0.0;
            for (int c = lane; c < a.ntg; c += 32) s0 += a.s_pdf[slab + (size_t)iu * a.ntg + c];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, off);
            if (lane == 0) sm.margu[iu] = s0 / A;                       // OTlib.py:93,156
        }
        for (int c = tid; c < a.ntg; c += 256) sm.margt[c] = sm.margt[c] / A;   // OTlib.py:93,155
        __syncthreads();

        // ---------------- P3: 1-D OT per marginal
        const size_t trow = a.tgt_per_window ? (size_t)b : 0;
        OtScratch sc{sm.cf, sm.tk, sm.dx, sm.E, sm.posf, sm.red};
        for (int c = tid; c < a.ntg; c += 256) sm.cf[c] = sm.margt[c];
        __syncthreads();
        const OtResult rt = block_ot1d(sc, a.ntg, a.tgt_cdf_t + trow * a.ntg, a.ntg, sm.xt,
                                       a.tgt_x_t + trow * a.ntg, a.pmask,
                                       (a.pmask & 1) ? sm.Rt : nullptr, (a.pmask & 2) ? sm.Rt : nullptr, nullptr);
        double gp = 0.0;
        for (int c = tid; c < a.ntg; c += 256) gp += sm.margt[c] * sm.Rt[c];
        const double Gt = block_sum(gp, sm.red);                        // <dwpmargX, pbar> (OTlib.py:1144)
        for (int c = tid; c < a.nug; c += 256) sm.cf[c] = sm.margu[c];
        __syncthreads();
        const OtResult ru = block_ot1d(sc, a.nug, a.tgt_cdf_u + trow * a.nug, a.nug, sm.xu,
                                       a.tgt_x_u + trow * a.nug, a.pmask,
                                       (a.pmask & 1) ? sm.Ru : nullptr, (a.pmask & 2) ? sm.Ru : nullptr, nullptr);
        gp = 0.0;
        for (int c = tid; c < a.nug; c += 256) gp += sm.margu[c] * sm.Ru[c];
        const double Gu = block_sum(gp, sm.red);                        // OTlib.py:1145
        common += (tid == 0) ? (rt.common + ru.common) : 0;
        if (tid == 0) {
            a.W[2 * (size_t)b] = (a.pmask & 1) ? rt.W1 : rt.W2;
            a.W[2 * (size_t)b + 1] = (a.pmask & 1) ? ru.W1 : ru.W2;
            if (a.dwg) a.dwg[b] = (a.pmask & 1) ? rt.dpos1 : rt.dpos2;  // OTlib.py:1121
        }

        // ---------------- P4: gradient assembly (FingerprintLib.py:205-228)
        if (a.grad) {
            for (int c = tid; c < a.ntg; c += 256) {
                const double ct = (sm.Rt[c] - Gt) / A;                   // dwpmargX (OTlib.py:1144,1146)
                int cur = -1;
                double t0 = 0.0, t1 = 0.0, u0 = 0.0, u1 = 0.0;
                for (int iu = 0; iu < a.nug; ++iu) {
                    const size_t k = slab + (size_t)iu * a.ntg + c;
                    const int i = a.s_idx[k];
                    if (i != cur) {
                        if (cur >= 0) {
                            atomicAdd(&sm.gbins[cur], t0); atomicAdd(&sm.gbins[cur + 1], t1);
                            atomicAdd(&sm.gbins[a.nt + cur], u0); atomicAdd(&sm.gbins[a.nt + cur + 1], u1);
                        }
                        cur = i; t0 = t1 = u0 = u1 = 0.0;
                    }
                    const double cu = (sm.Ru[iu] - Gu) / A;              // dwpmargY (OTlib.py:1145,1147)
                    const double wa = a.s_wa[k], wb = a.s_wb[k];
                    t0 += wa * ct; t1 += wb * ct; u0 += wa * cu; u1 += wb * cu;
                }
                if (cur >= 0) {
                    atomicAdd(&sm.gbins[cur], t0); atomicAdd(&sm.gbins[cur + 1], t1);
                    atomicAdd(&sm.gbins[a.nt + cur], u0); atomicAdd(&sm.gbins[a.nt + cur + 1], u1);
                }
            }
            __syncthreads();
            const double scale = -1.0 / (a.lambda * hdr.du);             // FingerprintLib.py:228,376-378
            for (int j = tid; j < a.nt; j += 256) {
                double chain = scale;
                if (a.transform) {   // d(un)/du, ricker_util.py:273,393-397
                    const double wj = load_sample(a.w, a.dtype, (long long)b * a.nt + j);
                    const double up = ((wj - hdr.u0raw) + (wj - hdr.u1raw)) / (hdr.u1raw - hdr.u0raw);
                    chain *= 2.0 / ((hdr.u1raw - hdr.u0raw) * CUDART_PI * (1.0 + up * up));
                }
                a.grad[((size_t)b * 2) * a.nt + j] = sm.gbins[j] * chain;
                a.grad[((size_t)b * 2 + 1) * a.nt + j] = sm.gbins[a.nt + j] * chain;
            }
        }
        __syncthreads();
    }
    if (a.status) {
        if (zero_dist) atomicAdd(a.status + WFOT_STAT_ZERO_DIST, zero_dist);
        if (slow) atomicAdd(a.status + WFOT_STAT_SLOW_PIXELS, slow);
        if (common) atomicAdd(a.status + WFOT_STAT_COMMON_CDF, common);
        if (degen) atomicAdd(a.status + WFOT_STAT_DEGENERATE_SEG, degen);
    }
}

static int fused_resident_ctas(size_t smem, int* per_sm_out) {
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_misfit_grad<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_misfit_grad<8>, 256, smem) != cudaSuccess) return -1;
    if (per_sm < 1) return -1;
    if (per_sm_out) *per_sm_out = per_sm;
    return sms * per_sm;
}

}  // namespace wfot

using namespace wfot;

extern "C" {

// Scratch: 28 bytes per pixel per resident CTA.  Sized for the largest grid the
// device can co-schedule (SM count x 4 CTAs) so the query needs no device call.
size_t wfot_misfit_grad_workspace_bytes(int B, int nt, int nug, int ntg) {
    if (B <= 0 || nt < 2 || nug < 1 || ntg < 1) return 0;
    int sms = wfot_device_sm_count();
    if (sms <= 0) sms = 148;
    size_t ctas = (size_t)sms * 4;
    if ((size_t)B < ctas) ctas = (size_t)B;
    return ctas * (size_t)nug * ntg * 28 + 256;
}

int wfot_misfit_grad_batch(const void* t, const void* w, int in_dtype, long long t_stride, int nt,
                           const wfot_grid* grids, int n_grids, int B, int nug, int ntg, double lambda,
                           int q, int pmask, int transform, const double* tgt_cdf_t, const double* tgt_x_t,
                           const double* tgt_cdf_u, const double* tgt_x_u, int tgt_per_window, double* W,
                           double* grad, double* dwg, void* workspace, size_t workspace_bytes,
                           int32_t* status, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!t || !w || !grids || !W || !workspace || !tgt_cdf_t || !tgt_x_t || !tgt_cdf_u || !tgt_x_u ||
        B <= 0 || nt < 2 || nug < 1 || ntg < 1 || (n_grids != 1 && n_grids != B) ||
        (q != 0 && q != 2) || (pmask != WFOT_W1 && pmask != WFOT_W2) || !(lambda > 0.0) ||
        (in_dtype != WFOT_F32 && in_dtype != WFOT_F64))
        return WFOT_ERR_INVALID_ARG;
    FusedArgs a;
    a.t = t; a.w = w; a.dtype = in_dtype; a.t_stride = t_stride; a.nt = nt; a.grids = grids;
    a.n_grids = n_grids; a.B = B; a.nug = nug; a.ntg = ntg; a.lambda = lambda; a.q = q; a.pmask = pmask;
    a.transform = transform; a.tgt_cdf_t = tgt_cdf_t; a.tgt_x_t = tgt_x_t; a.tgt_cdf_u = tgt_cdf_u;
    a.tgt_x_u = tgt_x_u; a.tgt_per_window = tgt_per_window; a.W = W; a.grad = grad; a.dwg = dwg;
    a.status = status;
    a.Spad = seg_pad(nt); a.ntg_pad = pad4(ntg); a.nug_pad = pad4(nug);
    a.nmax = pad4(ntg > nug ? ntg : nug);
    const size_t smem = fused_smem_bytes(nt, a.Spad, a.ntg_pad, a.nug_pad, a.nmax);
    if (smem > 227 * 1024) return WFOT_ERR_UNSUPPORTED;
    int per_sm = 0;
    int ctas = fused_resident_ctas(smem, &per_sm);
    if (ctas < 1) return cuda_fail(cudaGetLastError(), "k_misfit_grad occupancy");
    if (ctas > B) ctas = B;
    const size_t npix = (size_t)nug * ntg;
    uintptr_t base = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    const size_t avail = workspace_bytes - (base - (uintptr_t)workspace);
    const size_t max_ctas = avail / (npix * 28);
    if (max_ctas < 1) return WFOT_ERR_WORKSPACE;
    if ((size_t)ctas > max_ctas) ctas = (int)max_ctas;
    unsigned char* p = (unsigned char*)base;
    a.s_pdf = (double*)p;   p += (size_t)ctas * npix * 8;
    a.s_wa = (double*)p;    p += (size_t)ctas * npix * 8;
    a.s_wb = (double*)p;    p += (size_t)ctas * npix * 8;
    a.s_idx = (int32_t*)p;
    k_misfit_grad<8><<<ctas, 256, smem, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "wfot_misfit_grad_batch launch");
    return WFOT_OK;
}

}  // extern "C"
