// wfot_ot.cuh -- block-cooperative 1-D optimal transport (FP64):
// OTpdf normalisation + CDF prefix scan (libs/OTlib.py:92-93,112-114), stable
// rank-merge of the two monotone CDFs (libs/OTlib.py:668-673: append, argsort,
// bisect_left), W_1 / W_2^2 and translation derivatives (:690-706), and the
// derivative w.r.t. un-normalised source amplitudes in the O(n) suffix-sum form
// of the dense expression at :682-686,694,704.
// Every reduction/scan has a fixed order, so results are run-to-run identical.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wfot {

// ---- block-wide deterministic primitives (every thread of the block calls) ----
// `red` = 33 doubles of shared scratch.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < nw; ++i) r += red[i];
    return r;
}

// In-place inclusive scan of a[0..n) in shared memory.  reverse: suffix sums.
__device__ __forceinline__ void block_scan(double* a, int n, bool reverse, double* red) {
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int chunk = (n + T - 1) / T;
    const int beg = tid * chunk, end = min(beg + chunk, n);
    double run = 0.0;
    for (int i = beg; i < end; ++i) {
        const int j = reverse ? n - 1 - i : i;
        run += a[j];
        a[j] = run;
    }
    // exclusive scan of the per-thread totals
    double inc = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
    }
    __syncthreads();
    if (lane == 31) red[wid] = inc;
    __syncthreads();
    double woff = 0.0;
    for (int i = 0; i < wid; ++i) woff += red[i];
    const double excl = woff + inc - run;
    for (int i = beg; i < end; ++i) {
        const int j = reverse ? n - 1 - i : i;
        a[j] += excl;
    }
    __syncthreads();
}

// In-place inclusive running maximum of a[0..n) in shared memory (exact: max is associative).
__device__ __forceinline__ void block_running_max(double* a, int n, double* red) {
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int chunk = (n + T - 1) / T;
    const int beg = tid * chunk, end = min(beg + chunk, n);
    const double ninf = __longlong_as_double(0xfff0000000000000LL);
    double run = ninf;
    for (int i = beg; i < end; ++i) { run = fmax(run, a[i]); a[i] = run; }
    double inc = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc = fmax(inc, o);
    }
    double excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = ninf;
    __syncthreads();
    if (lane == 31) red[wid] = inc;
    __syncthreads();
    for (int i = 0; i < wid; ++i) excl = fmax(excl, red[i]);
    for (int i = beg; i < end; ++i) a[i] = fmax(a[i], excl);
    __syncthreads();
}

// np.cumsum of non-negative terms (libs/OTlib.py:113) for a[0..n) in shared memory; returns cumsum[-1].
// The reference's sequential sum is non-decreasing by construction.  A parallel scan groups the additions
// differently, and where a term is smaller than the ulp of the running sum (density tails for small lambda)
// a later prefix can round BELOW an earlier one.  The CDF merge needs sorted knots, so such entries are raised
// to the running maximum (a change of at most the scan's own rounding error).  No extra barrier when the scan
// is already monotone.
__device__ __forceinline__ double block_cumsum(double* a, int n, double* red) {
    block_scan(a, n, false, red);
    int viol = 0;
    for (int j = threadIdx.x + 1; j < n; j += blockDim.x) viol |= (a[j] < a[j - 1]);
    double last = a[n - 1];
    if (__syncthreads_or(viol)) {
        block_running_max(a, n, red);
        last = a[n - 1];
        __syncthreads();
    }
    return last;
}

// ---- canonical (CTA-shape independent) reductions -------------------------------------------------------
// The fused path compares the CDFs of a predicted window with those of the observed window for EXACT equality
// (libs/OTlib.py:663-666 raises TargetSourceCDFError on common values; identical windows are the practical
// trigger).  Observed and predicted CDFs are produced by different launches - different CTA sizes, one kernel
// or two - so every sum that feeds a CDF bit uses an order that does not depend on the launch shape: warp 0
// alone, lane k adding elements k, k + 32, ... in ascending order, then a xor-shuffle tree; the prefix sum of
// canon_cdf() is sequential (np.cumsum's own order).  n is a grid axis (tens to ~1000 entries).
// Every thread of the block calls; `red` = 2 doubles of shared scratch.
__device__ __forceinline__ double canon_sum(const double* a, int n, double* red) {
    if (threadIdx.x < 32) {
        double v = 0.0;
        for (int j = threadIdx.x; j < n; j += 32) v += a[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    const double r = red[0];
    __syncthreads();
    return r;
}

// OTpdf of a 1-D density in shared memory (libs/OTlib.py:91-93,112-114):
// a[0..n) un-normalised amplitudes -> CDF = cumsum(a / amp) / cumsum(a / amp)[-1].  Returns amp; *neg (nullable,
// valid in every thread) = number of negative amplitudes.
// The prefix sum is SEQUENTIAL, one lane, exactly np.cumsum's order (:113).  That is affordable here - n is a grid
// axis, the chain is n dependent additions (~2 k cycles for 256 bins against ~10^6 per window) - and it buys the
// reference's behaviour where a parallel scan only approximates it: the result is non-decreasing by construction,
// and a density tail below the ulp of the running sum saturates to EXACTLY the total, so both CDFs of such a pair
// hold 1.0 before their last entry and the reference's common-CDF check (:663-666) fires - here as there.
__device__ __forceinline__ double canon_cdf(double* a, int n, double* red, int* neg) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        double v = 0.0;
        int ng = 0;
        for (int j = lane; j < n; j += 32) { const double x = a[j]; v += x; ng += (x < 0.0); }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, off);
            ng += __shfl_xor_sync(0xffffffffu, ng, off);
        }
        const double amp = v;
        for (int j = lane; j < n; j += 32) a[j] = a[j] / amp;    // pdf / amp (:93)
        __syncwarp();
        if (lane == 0) {
            double run = 0.0;
            for (int j = 0; j < n; ++j) { run += a[j]; a[j] = run; }
        }
        __syncwarp();
        const double last = a[n - 1];
        __syncwarp();
        for (int j = lane; j < n; j += 32) a[j] = a[j] / last;   // cdf / cdf[-1] (:114)
        if (lane == 0) { red[0] = amp; red[1] = __longlong_as_double((long long)ng); }
    }
    __syncthreads();
    const double amp = red[0];
    if (neg) *neg = (int)__double_as_longlong(red[1]);
    __syncthreads();
    return amp;
}

__device__ __forceinline__ int lower_bound_d(const double* a, int n, double v) {   // bisect_left
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int upper_bound_d(const double* a, int n, double v) {   // bisect_right
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Shared-memory working set of one 1-D problem.
struct OtScratch {
    double* cf;    // [n]   in: un-normalised source amplitudes; out: source CDF
    double* tk;    // [n+m-1] merged knots
    double* dx;    // [n+m-1] x_f[indf] - x_g[indg]
    double* E;     // [n]   derivative work array; out: dW/df (last order computed)
    int* posf;     // [n]   merged position of source knot j
    double* red;   // [33]
};

struct OtResult {
    double amp;        // sum of the un-normalised source amplitudes
    double W1, W2;     // W_1 and W_2^2
    double dpos1, dpos2;
    int neg, common;
};

// cg/xg/xf may live in shared or global memory.  If dW1/dW2 != nullptr the
// derivative vectors (length n) are written there (any address space).
// merge_order (nullable, global): the reference's tkarg.
__device__ __forceinline__ OtResult block_ot1d(const OtScratch& sc, int n, const double* cg, int m,
                                               const double* xf, const double* xg, int pmask,
                                               double* dW1, double* dW2, int32_t* merge_order) {
    const int tid = threadIdx.x, T = blockDim.x;
    OtResult r;
    // -- OTpdf: sign check, normalise, CDF (libs/OTlib.py:91-93,112-114), canonical summation order
    int neg = 0;
    r.amp = canon_cdf(sc.cf, n, sc.red, &neg);
    // -- stable rank-merge of cf[:-1] and cg (:668-672)
    const int K = n + m - 1;
    int common = 0;
    for (int j = tid; j < n - 1; j += T) {
        const double v = sc.cf[j];
        const int lb = lower_bound_d(cg, m, v);
        const int pos = j + lb;                       // source knots first on ties
        const int a = lower_bound_d(sc.cf, n, v);     // bisect_left(cf, tk)
        sc.tk[pos] = v;
        sc.dx[pos] = xf[a] - xg[lb];
        sc.posf[j] = pos;
        if (lb < m - 1 && cg[lb] == v) ++common;      // np.intersect1d(cg[:-1], cf[:-1]) (:664)
        if (merge_order) merge_order[pos] = j;
    }
    for (int i = tid; i < m; i += T) {
        const double v = cg[i];
        const int ub = upper_bound_d(sc.cf, n - 1, v);
        const int pos = i + ub;
        const int a = lower_bound_d(sc.cf, n, v);
        const int b = lower_bound_d(cg, m, v);
        sc.tk[pos] = v;
        sc.dx[pos] = xf[a] - xg[b];
        if (merge_order) merge_order[pos] = n - 1 + i;
    }
    __syncthreads();
    // -- W_p^p and translation derivatives (:673,690-706)
    double w1 = 0.0, w2 = 0.0, p1 = 0.0, p2 = 0.0;
    for (int k = tid; k < K; k += T) {
        const double dt = k ? sc.tk[k] - sc.tk[k - 1] : sc.tk[0];
        const double d = sc.dx[k];
        w1 += fabs(d) * dt;
        w2 += d * d * dt;
        p1 += (d > 0.0 ? dt : (d < 0.0 ? -dt : 0.0));
        p2 += 2.0 * d * dt;
    }
    r.W1 = block_sum(w1, sc.red);
    r.W2 = block_sum(w2, sc.red);
    r.dpos1 = block_sum(p1, sc.red);
    r.dpos2 = block_sum(p2, sc.red);
    r.neg = neg;
    r.common = (int)block_sum((double)common, sc.red);
    // -- d/d(un-normalised source amplitude) (:682-686,694,704), O(n) form
    for (int p = 1; p <= 2; ++p) {
        double* out = (p == 1) ? dW1 : dW2;
        if (!(pmask & p) || out == nullptr) continue;
        double z = 0.0;
        for (int j = tid; j < n; j += T) {
            double e = 0.0;
            if (j < n - 1) {
                const int k = sc.posf[j];
                const double d0 = sc.dx[k];
                const double d1 = (k + 1 < K) ? sc.dx[k + 1] : 0.0;
                const double c0 = (p == 1) ? fabs(d0) : d0 * d0;
                const double c1 = (k + 1 < K) ? ((p == 1) ? fabs(d1) : d1 * d1) : 0.0;
                e = c0 - c1;
            }
            sc.E[j] = e;
            z += sc.cf[j] * e;
        }
        const double Z = block_sum(z, sc.red);
        block_scan(sc.E, n, true, sc.red);
        for (int j = tid; j < n; j += T) out[j] = (sc.E[j] - Z) / r.amp;
        __syncthreads();
    }
    return r;
}

}  // namespace wfot
