// wfot_ot.cuh -- 1-D optimal transport inside the fused path (FP64), one warp per problem:
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

// ------------------------------------------------------------------ one warp, no block barriers
// The fused path solves two small problems per window (n = m = a grid axis, tens to a few hundred bins).  The
// block-cooperative version of round 1 ended every phase in a block barrier or a block reduction, and a window paid
// their latency, not the work: ~24 k cycles per problem whether it had 61 or 256 bins (29 % of a 79 x 61 window).
// warp_ot1d() is the same algorithm run by ONE warp with warp-level synchronisation only - the two marginal problems
// of a window run on two warps side by side when shared memory has room for two scratch sets - the rank searches are
// branch-free with a fixed trip count and U elements of a lane are searched side by side, so their shared-memory
// latencies overlap.
// Same CDF (canonical order: the arithmetic of canon_cdf()); the sums over the merged knots are lane-strided + xor tree.

// bisect_left (LEQ = false) / bisect_right (LEQ = true) of U values at once in a[0..n), n >= 1, warp-uniform trip count
template <int U, bool LEQ>
__device__ __forceinline__ void bisect_multi(const double* a, int n, const double (&v)[U], int (&lo)[U]) {
#pragma unroll
    for (int u = 0; u < U; ++u) lo[u] = 0;
    int len = n;
    while (len > 1) {
        const int half = len >> 1;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double x = a[lo[u] + half - 1];
            lo[u] += (LEQ ? (x <= v[u]) : (x < v[u])) ? half : 0;
        }
        len -= half;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const double x = a[lo[u]];
        lo[u] += (LEQ ? (x <= v[u]) : (x < v[u])) ? 1 : 0;
    }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Called by all 32 lanes of one warp.  sc.cf: un-normalised source amplitudes (n) on entry, source CDF on return;
// cg_g / xg_g: target CDF / abscissae in global memory (m = n entries); xf: source abscissae (shared); dW (shared, n
// doubles): receives d W_p^p / d(un-normalised source amplitude) for the order(s) in pmask (the higher one last).
// `marg` (shared, n): the normalised source density; the result carries G = <dW, marg> (OTlib.py:1144-1145).
struct WarpOtResult { OtResult r; double G; };

__device__ __forceinline__ WarpOtResult warp_ot1d(const OtScratch& sc, int n, const double* cg_g, const double* xf,
                                                  const double* xg_g, int pmask, double* dW, const double* marg) {
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const int m = n, K = 2 * n - 1;
    WarpOtResult out;
    double* const cf = sc.cf;
    double* const cg = sc.E;      // target CDF staged for the searches (E is written after the merge) ...
    double* const xg = dW;        // ... and its abscissae (dW likewise)
    for (int j = lane; j < m; j += 32) { cg[j] = cg_g[j]; xg[j] = xg_g[j]; }
    // -- OTpdf: sign check, normalise, CDF (libs/OTlib.py:91-93,112-114): the arithmetic and order of canon_cdf()
    {
        double v = 0.0;
        int ng = 0;
        for (int j = lane; j < n; j += 32) { const double x = cf[j]; v += x; ng += (x < 0.0); }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, off);
            ng += __shfl_xor_sync(0xffffffffu, ng, off);
        }
        out.r.amp = v; out.r.neg = ng;
        for (int j = lane; j < n; j += 32) cf[j] = cf[j] / v;      // pdf / amp (:93)
        __syncwarp();
        if (lane == 0) {
            double run = 0.0;
            for (int j = 0; j < n; ++j) { run += cf[j]; cf[j] = run; }
        }
        __syncwarp();
        const double last = cf[n - 1];
        __syncwarp();
        for (int j = lane; j < n; j += 32) cf[j] = cf[j] / last;   // cdf / cdf[-1] (:114)
        __syncwarp();
    }
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    // -- stable rank-merge of cf[:-1] and cg (:668-672)
    int common = 0;
    for (int j0 = lane; j0 < n - 1; j0 += 32 * U) {
        double v[U];
        int lb[U], a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (j0 + 32 * u < n - 1) ? cf[j0 + 32 * u] : kInf;
        bisect_multi<U, false>(cg, m, v, lb);
        bisect_multi<U, false>(cf, n, v, a);              // bisect_left(cf, tk)
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = j0 + 32 * u;
            if (j < n - 1) {
                const int pos = j + lb[u];                // source knots first on ties
                sc.tk[pos] = v[u];
                sc.dx[pos] = xf[a[u]] - xg[lb[u]];
                sc.posf[j] = pos;
                if (lb[u] < m - 1 && cg[lb[u]] == v[u]) ++common;      // np.intersect1d(cg[:-1], cf[:-1]) (:664)
            }
        }
    }
    for (int i0 = lane; i0 < m; i0 += 32 * U) {
        double v[U];
        int ub[U], a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (i0 + 32 * u < m) ? cg[i0 + 32 * u] : kInf;
        if (n > 1) bisect_multi<U, true>(cf, n - 1, v, ub);
        else {
#pragma unroll
            for (int u = 0; u < U; ++u) ub[u] = 0;
        }
        bisect_multi<U, false>(cf, n, v, a);
        bisect_multi<U, false>(cg, m, v, b);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + 32 * u;
            if (i < m) {
                const int pos = i + ub[u];
                sc.tk[pos] = v[u];
                sc.dx[pos] = xf[a[u]] - xg[b[u]];
            }
        }
    }
    __syncwarp();
    // -- W_p^p and translation derivatives (:673,690-706)
    double w1 = 0.0, w2 = 0.0, p1 = 0.0, p2 = 0.0;
    for (int k = lane; k < K; k += 32) {
        const double dt = k ? sc.tk[k] - sc.tk[k - 1] : sc.tk[0];
        const double d = sc.dx[k];
        w1 += fabs(d) * dt;
        w2 += d * d * dt;
        p1 += (d > 0.0 ? dt : (d < 0.0 ? -dt : 0.0));
        p2 += 2.0 * d * dt;
    }
    out.r.W1 = warp_sum_d(w1);
    out.r.W2 = warp_sum_d(w2);
    out.r.dpos1 = warp_sum_d(p1);
    out.r.dpos2 = warp_sum_d(p2);
    out.r.common = __reduce_add_sync(0xffffffffu, common);
    __syncwarp();                                     // cg (in E) and xg (in dW) are dead from here on
    // -- d/d(un-normalised source amplitude) (:682-686,694,704), O(n) form
    const int chunk = (n + 31) >> 5;
    const int beg = lane * chunk, end = min(beg + chunk, n);
    for (int p = 1; p <= 2; ++p) {
        if (!(pmask & p)) continue;
        double z = 0.0;
        for (int j = lane; j < n; j += 32) {
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
            z += cf[j] * e;
        }
        const double Z = warp_sum_d(z);
        __syncwarp();
        // suffix sums of E: a lane owns a contiguous chunk (walked from the end), lane totals combined by a warp scan
        double run = 0.0;
        for (int i = beg; i < end; ++i) { const int j = n - 1 - i; run += sc.E[j]; sc.E[j] = run; }
        double inc = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double o = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += o;
        }
        const double excl = inc - run;
        for (int i = beg; i < end; ++i) sc.E[n - 1 - i] += excl;
        __syncwarp();
        for (int j = lane; j < n; j += 32) dW[j] = (sc.E[j] - Z) / out.r.amp;
        __syncwarp();
    }
    double g = 0.0;
    for (int j = lane; j < n; j += 32) g += marg[j] * dW[j];
    out.G = warp_sum_d(g);
    return out;
}

}  // namespace wfot
