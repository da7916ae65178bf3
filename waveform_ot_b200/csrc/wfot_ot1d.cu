// wfot_ot1d.cu -- batched 1-D optimal transport, one WARP per (source, target) pair (sm_100a).
//
// Replaces OTpdf.__init__ (1-D) + wasser(distfunc in {'W1','W2','W12'}, derivatives=...):
// libs/OTlib.py:90-117 (normalise, cumsum -> CDF) and :596-706 (merge, W_p^p, derivatives).
//
// A warp owns one pair at a time; nothing but warp-level synchronisation is used:
//   1. load f / g with 16-byte coalesced loads (lane L owns elements 4L..4L+3 of every 128-element
//      round), FP64 sum -> amp, pdf/amp, FP64 prefix scan (4-element local scan + one warp scan per
//      round), /last -> the two CDFs in shared memory (libs/OTlib.py:92-93,112-114);
//   2. merge path: lane d binary-searches where diagonal d*ceil(K/32) of the (cf[:-1], cg) merge grid
//      is crossed (source first on ties = a stable argsort of the concatenation, :668-669);
//   3. each lane merges its K/32 knots sequentially: quantile ranks (bisect_left semantics, :671-672,
//      including repeated CDF values), dx = x_f[indf] - x_g[indg], dt, and the fused epilogue
//      W1 += |dx| dt, W2 += dx^2 dt, d/dx0 (:690-706); for the amplitude derivative the O(n) form of
//      the dense (n x (n+m-1)) product at :682-686,694,704:
//      E_j = |dx|^p at the merged position of source knot j minus the next knot's;
//   4. suffix scan of E_j (parked in the dW output rows themselves) -> dW_i = (sum_{j>=i} E_j - Z)/amp.
// Divisions by the per-pair scalars (amp, cdf[-1]) use one correctly rounded reciprocal and a fused
// residual correction (q = x r; q += fma(-d, q, x) r), which reproduces IEEE division except in rare
// last-bit cases; everything else is plain FP64 in a fixed order (run-to-run identical results).
//
// Algorithmic HBM bytes per pair: 4 (n + m) in (FP32 input) + 8 n per derivative vector out.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/wfot.h"
#include "wfot_host.h"

namespace wfot {

struct Ot1dArgs {
    const void* f; const void* g; int dtype; const double* xf; const double* xg;
    long long f_stride, g_stride, xf_stride, xg_stride; int n, m, pmask, deriv; long long B;
    double* W; double* dW1; double* dW2; double* dpos; double* amp_f; double* cdf_f; double* cdf_g;
    int32_t* merge_order; int32_t* status;
    int npad, mpad, wpc, xshared, need_e2;
    size_t per_warp;
};

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
    return v;
}

// explicit shared-space accesses with 32-bit addresses (the merge loop's pointers otherwise drag generic ->
// shared address arithmetic through every step)
__device__ __forceinline__ double lds64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// x / d with r = RN(1/d): one multiply and a fused residual correction
__device__ __forceinline__ double div_by(double x, double d, double r) {
    const double q = x * r;
    return fma(fma(-d, q, x), r, q);
}

// four consecutive samples idx..idx+3 of a row as doubles (0 beyond n); 16-byte loads when aligned
__device__ __forceinline__ void load4(const void* row, int dtype, bool vec, int idx, int n, double (&v)[4]) {
    if (vec && idx + 3 < n) {
        if (dtype == WFOT_F32) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + idx));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
            const double2* p = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(row) + idx);
            const double2 q0 = __ldg(p), q1 = __ldg(p + 1);
            v[0] = q0.x; v[1] = q0.y; v[2] = q1.x; v[3] = q1.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double s = 0.0;
            if (idx + i < n)
                s = dtype == WFOT_F32 ? (double)reinterpret_cast<const float*>(row)[idx + i]
                                      : reinterpret_cast<const double*>(row)[idx + i];
            v[i] = s;
        }
    }
}

// Running maximum over c[0..npad): a CDF that is not strictly increasing may hold prefix sums that the parallel
// scan rounded out of order (terms below the ulp of the running sum; see block_cumsum in wfot_ot.cuh).  The
// reference's sequential np.cumsum is non-decreasing by construction and the merge needs sorted knots, so such
// entries of the (already rescaled) CDF are raised to the running maximum.  Only reached by non-strict CDFs
// (rare), a no-op on sorted ones.
__device__ __forceinline__ void warp_running_max(double* c, int n, int npad, double* cdf_out, int lane) {
    double carry = -CUDART_INF;
    for (int base = 0; base < npad; base += 128) {
        const int idx = base + 4 * lane;
        const double2 a0 = *reinterpret_cast<const double2*>(c + idx);
        const double2 a1 = *reinterpret_cast<const double2*>(c + idx + 2);
        double v0 = a0.x, v1 = fmax(v0, a0.y), v2 = fmax(v1, a1.x), v3 = fmax(v2, a1.y);
        double inc = v3;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double o = __shfl_up_sync(kFull, inc, off);
            if (lane >= off) inc = fmax(inc, o);
        }
        double excl = __shfl_up_sync(kFull, inc, 1);
        excl = lane == 0 ? carry : fmax(excl, carry);
        // the rescaled CDF ends at exactly 1: an earlier prefix that rounded above the last one is capped there
        v0 = fmin(fmax(v0, excl), 1.0); v1 = fmin(fmax(v1, excl), 1.0);
        v2 = fmin(fmax(v2, excl), 1.0); v3 = fmin(fmax(v3, excl), 1.0);
        *reinterpret_cast<double2*>(c + idx) = make_double2(v0, v1);
        *reinterpret_cast<double2*>(c + idx + 2) = make_double2(v2, v3);
        if (cdf_out) {
            if (idx < n) cdf_out[idx] = v0;
            if (idx + 1 < n) cdf_out[idx + 1] = v1;
            if (idx + 2 < n) cdf_out[idx + 2] = v2;
            if (idx + 3 < n) cdf_out[idx + 3] = v3;
        }
        carry = fmax(carry, __shfl_sync(kFull, inc, 31));
    }
    __syncwarp();
}

// Register-resident OTpdf.__init__ for FP32 rows of n <= 1024 (16-byte aligned): the samples stay in
// registers from the global load to the final CDF value, which is written to shared memory once
// (the generic path below makes three shared-memory round trips).  Same arithmetic, same order.
__device__ __forceinline__ double warp_cdf_f32_1k(const float* fr, int n, int npad, double* c, double* cdf_out,
                                                  int lane, int& neg, bool& strict) {
    float4 q[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int idx = r * 128 + 4 * lane;
        q[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx + 3 < n) q[r] = __ldg(reinterpret_cast<const float4*>(fr + idx));
        else if (idx < n) {
            q[r].x = fr[idx];
            if (idx + 1 < n) q[r].y = fr[idx + 1];
            if (idx + 2 < n) q[r].z = fr[idx + 2];
        }
    }
    double s = 0.0;
    float vmin = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        vmin = fminf(vmin, fminf(fminf(q[r].x, q[r].y), fminf(q[r].z, q[r].w)));
        s += ((double)q[r].x + (double)q[r].y) + ((double)q[r].z + (double)q[r].w);
    }
    neg |= (vmin < 0.f);
    const double amp = warp_sum(s);                                   // :92
    const double ramp = 1.0 / amp;
    double cv[8][4];
    double carry = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if (r * 128 < npad) {
            const double c0 = div_by((double)q[r].x, amp, ramp);      // pdf / amp (:93)
            const double c1 = c0 + div_by((double)q[r].y, amp, ramp);
            const double c2 = c1 + div_by((double)q[r].z, amp, ramp);
            const double c3 = c2 + div_by((double)q[r].w, amp, ramp);
            double inc = c3;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const double o = __shfl_up_sync(kFull, inc, off);
                if (lane >= off) inc += o;
            }
            const double add = carry + (inc - c3);
            cv[r][0] = add + c0; cv[r][1] = add + c1; cv[r][2] = add + c2; cv[r][3] = add + c3;
            carry += __shfl_sync(kFull, inc, 31);
        } else {
            cv[r][0] = cv[r][1] = cv[r][2] = cv[r][3] = 0.0;
        }
    }
    // cumsum[-1] (:113): the value held for element n-1
    const int rn = (n - 1) >> 7, ln = ((n - 1) & 127) >> 2, sn = (n - 1) & 3;
    double mine = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (r == rn && i == sn) mine = cv[r][i];
    const double last = __shfl_sync(kFull, mine, ln);
    const double rl = 1.0 / last;
    const bool resc = (last != 1.0);
    bool ok = true;
    double prev_hi = -CUDART_INF;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if (r * 128 < npad) {
            const int idx = r * 128 + 4 * lane;
            double a0 = cv[r][0], a1 = cv[r][1], a2 = cv[r][2], a3 = cv[r][3];
            if (resc) {                                               // cdf / cdf[-1] (:114)
                a0 = div_by(a0, last, rl); a1 = div_by(a1, last, rl);
                a2 = div_by(a2, last, rl); a3 = div_by(a3, last, rl);
            }
            *reinterpret_cast<double2*>(c + idx) = make_double2(a0, a1);
            *reinterpret_cast<double2*>(c + idx + 2) = make_double2(a2, a3);
            double left = __shfl_up_sync(kFull, a3, 1);
            if (lane == 0) left = prev_hi;
            prev_hi = __shfl_sync(kFull, a3, 31);
            ok = ok && (idx >= n || left < a0) && (idx + 1 >= n || a0 < a1) &&
                 (idx + 2 >= n || a1 < a2) && (idx + 3 >= n || a2 < a3);
            if (cdf_out) {
                if (idx < n) cdf_out[idx] = a0;
                if (idx + 1 < n) cdf_out[idx + 1] = a1;
                if (idx + 2 < n) cdf_out[idx + 2] = a2;
                if (idx + 3 < n) cdf_out[idx + 3] = a3;
            }
        }
    }
    strict = __all_sync(kFull, ok);
    __syncwarp();
    if (!strict) warp_running_max(c, n, npad, cdf_out, lane);
    return amp;
}

// OTpdf.__init__ for one 1-D density (libs/OTlib.py:91-93,112-114): c[0..n) <- cdf, returns amp.
// c has room for npad = ceil(n/128)*128 doubles.  strict <- the CDF is strictly increasing.
__device__ __forceinline__ double warp_cdf(const void* row, int dtype, int n, int npad, double* c,
                                           double* cdf_out, int lane, int& neg, bool& strict) {
    const bool vec = (reinterpret_cast<uintptr_t>(row) & 15) == 0;     // 16-byte loads need an aligned row
    if (dtype == WFOT_F32 && vec && npad <= 1024)
        return warp_cdf_f32_1k(reinterpret_cast<const float*>(row), n, npad, c, cdf_out, lane, neg, strict);
    double s = 0.0;
    double vmin = 0.0;
    if (dtype == WFOT_F32 && vec) {                                   // 8 x 16-byte loads in flight per lane
        const float* fr = reinterpret_cast<const float*>(row);
        for (int base0 = 0; base0 < npad; base0 += 8 * 128) {
            float4 q[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int idx = base0 + r * 128 + 4 * lane;
                q[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx + 3 < n) q[r] = __ldg(reinterpret_cast<const float4*>(fr + idx));
                else if (idx < n) {
                    q[r].x = fr[idx];
                    if (idx + 1 < n) q[r].y = fr[idx + 1];
                    if (idx + 2 < n) q[r].z = fr[idx + 2];
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int idx = base0 + r * 128 + 4 * lane;
                if (idx < npad) {
                    const double v0 = q[r].x, v1 = q[r].y, v2 = q[r].z, v3 = q[r].w;
                    vmin = fmin(vmin, (double)fminf(fminf(q[r].x, q[r].y), fminf(q[r].z, q[r].w)));
                    s += (v0 + v1) + (v2 + v3);
                    *reinterpret_cast<double2*>(c + idx) = make_double2(v0, v1);
                    *reinterpret_cast<double2*>(c + idx + 2) = make_double2(v2, v3);
                }
            }
        }
    } else {
        for (int base = 0; base < npad; base += 128) {
            const int idx = base + 4 * lane;
            double v[4];
            load4(row, dtype, vec, idx, n, v);
            vmin = fmin(vmin, fmin(fmin(v[0], v[1]), fmin(v[2], v[3])));
            s += (v[0] + v[1]) + (v[2] + v[3]);
            *reinterpret_cast<double2*>(c + idx) = make_double2(v[0], v[1]);
            *reinterpret_cast<double2*>(c + idx + 2) = make_double2(v[2], v[3]);
        }
    }
    neg |= (vmin < 0.0);
    const double amp = warp_sum(s);                                   // :92
    const double ramp = 1.0 / amp;
    double carry = 0.0;
    for (int base = 0; base < npad; base += 128) {
        const int idx = base + 4 * lane;
        const double2 a0 = *reinterpret_cast<const double2*>(c + idx);
        const double2 a1 = *reinterpret_cast<const double2*>(c + idx + 2);
        const double c0 = div_by(a0.x, amp, ramp);                    // pdf / amp (:93)
        const double c1 = c0 + div_by(a0.y, amp, ramp);
        const double c2 = c1 + div_by(a1.x, amp, ramp);
        const double c3 = c2 + div_by(a1.y, amp, ramp);
        double inc = c3;                                              // warp inclusive scan of the lane totals
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double o = __shfl_up_sync(kFull, inc, off);
            if (lane >= off) inc += o;
        }
        const double add = carry + (inc - c3);
        *reinterpret_cast<double2*>(c + idx) = make_double2(add + c0, add + c1);
        *reinterpret_cast<double2*>(c + idx + 2) = make_double2(add + c2, add + c3);
        carry += __shfl_sync(kFull, inc, 31);
    }
    __syncwarp();
    const double last = c[n - 1];                                     // cumsum[-1] (:113)
    __syncwarp();                                                     // every lane has read it before it is rescaled
    const double rl = 1.0 / last;
    const bool resc = (last != 1.0);
    bool ok = true;
    double prev_hi = -CUDART_INF;                                     // last element of the previous round
    for (int base = 0; base < npad; base += 128) {
        const int idx = base + 4 * lane;
        double2 a0 = *reinterpret_cast<const double2*>(c + idx);
        double2 a1 = *reinterpret_cast<const double2*>(c + idx + 2);
        if (resc) {                                                   // cdf / cdf[-1] (:114)
            a0.x = div_by(a0.x, last, rl); a0.y = div_by(a0.y, last, rl);
            a1.x = div_by(a1.x, last, rl); a1.y = div_by(a1.y, last, rl);
            *reinterpret_cast<double2*>(c + idx) = a0;
            *reinterpret_cast<double2*>(c + idx + 2) = a1;
        }
        double left = __shfl_up_sync(kFull, a1.y, 1);
        if (lane == 0) left = prev_hi;
        prev_hi = __shfl_sync(kFull, a1.y, 31);
        ok = ok && (idx >= n || left < a0.x) && (idx + 1 >= n || a0.x < a0.y) &&
             (idx + 2 >= n || a0.y < a1.x) && (idx + 3 >= n || a1.x < a1.y);
        if (cdf_out) {
            if (idx < n) cdf_out[idx] = a0.x;
            if (idx + 1 < n) cdf_out[idx + 1] = a0.y;
            if (idx + 2 < n) cdf_out[idx + 2] = a1.x;
            if (idx + 3 < n) cdf_out[idx + 3] = a1.y;
        }
    }
    strict = __all_sync(kFull, ok);
    __syncwarp();
    if (!strict) warp_running_max(c, n, npad, cdf_out, lane);
    return amp;
}

// ------------------------------------------------------------------ the merge
// Per-lane accumulators shared by the lane's two chains.
struct Acc {
    double w1, w2, p1, p2, z1, z2;
    int common;
};

// One contiguous share ("chain") of the merged knot sequence: source knots [ia, ia1), target knots
// [ib, ib1).  A lane runs TWO chains in lock step, which gives the scheduler two independent
// dependency chains per warp (the loop is latency bound: compare -> select -> rank -> x look-up).
//   STRICT: both CDFs strictly increasing -> the bisect_left ranks (:671-672) are the running counters
//           (minus one for a target knot equal to the last consumed source knot);
//   otherwise equal-value runs are tracked (repeated CDF values = empty bins).
//   E1 / E2: the per-source-knot E_j^{(p)} of the amplitude derivative are parked in shared memory,
//           slot j of e1s / e2s.  One of them is the cf array itself: entry j is dead once consumed and
//           a chain never reads outside its own ranges after the __syncwarp that follows init().
// Template flags: PM = orders whose W_p^p / translation derivative are accumulated (1, 2 or 3);
// EM = orders whose amplitude derivative is wanted (subset of PM); MO = write the merge order.
template <bool STRICT, int PM, int EM, bool MO>
struct Chain {
    static constexpr bool E1 = (EM & 1) != 0, E2 = (EM & 2) != 0;
    int ia, ia1, ib, ib1, k, pj;
    double va, vb, tprev, runf_val, rung_val;
    int runf_len, rung_len;
    double first_c1, first_c2, pc1, pc2, pcf;
    bool first;
    double xa, xb;                                // STRICT: x_f[ia], x_g[ib] travel with the heads

    // cf and cg (and x_f, x_g) are contiguous: cg = cf + npad, xg = xf + npad.
    __device__ __forceinline__ void init(const double* cf, const double* xf, int npad, int k0,
                                         int ia_, int ia1_, int ib_, int ib1_, int n, int m) {
        const double* cg = cf + npad;
        ia = ia_; ia1 = ia1_; ib = ib_; ib1 = ib1_; k = k0; pj = -1;
        tprev = 0.0; runf_val = CUDART_NAN; rung_val = CUDART_NAN; runf_len = 0; rung_len = 0;
        first_c1 = 0.0; first_c2 = 0.0; pc1 = 0.0; pc2 = 0.0; pcf = 0.0; first = true;
        if (ia > 0) {
            runf_val = cf[ia - 1]; runf_len = 1;
            if (!STRICT) while (ia - 1 - runf_len >= 0 && cf[ia - 1 - runf_len] == runf_val) ++runf_len;
            tprev = runf_val;
        }
        if (ib > 0) {
            rung_val = cg[ib - 1]; rung_len = 1;
            if (!STRICT) while (ib - 1 - rung_len >= 0 && cg[ib - 1 - rung_len] == rung_val) ++rung_len;
            tprev = ia > 0 ? fmax(tprev, rung_val) : rung_val;
        }
        va = ia < ia1 ? cf[ia] : CUDART_INF;      // heads, bounded by the chain's own ranges
        vb = ib < ib1 ? cg[ib] : CUDART_INF;
        xa = 0.0; xb = 0.0;
        if (STRICT) {                             // an empty trailing chain starts at ia = n-1, ib = m: clamp as step() does
            xa = xf[min(ia, n - 1)]; xb = xf[npad + min(ib, m - 1)];
            if ((E1 || E2) && (ia < ia1 || ib < ib1)) {      // |dx|^p of the chain's first knot (for the previous chain)
                const bool src = (va <= vb);
                const bool tie = !src && ia > 0 && (vb == tprev);
                const double dx = (tie ? xf[ia - 1] : xa) - xb;
                first_c1 = fabs(dx); first_c2 = dx * dx;
            }
        }
    }

    // One merged knot, branch-free (selects and predicated loads / stores only).
    // STRICT: bisect_left(cf, v) is ia (ia-1 for a target knot equal to the last consumed source knot,
    // which is then the previous knot: v == tprev) and bisect_left(cg, v) is ib, i.e. x_f[indf], x_g[indg]
    // are the x of the two current heads, which travel with them in registers; the consumed side's next
    // (cdf, x) pair is fetched with one selected index.
    // cfa / xfa / e1a / e2a: 32-bit shared-memory byte addresses of [cf | cg], [x_f | x_g] and the E arrays.
    __device__ __forceinline__ void step(uint32_t cfa, uint32_t xfa, int npad,
                                         uint32_t e1a, uint32_t e2a, int32_t* mo, int n, int m, Acc& acc) {
        const bool src = (va <= vb);              // source first on ties (stable argsort of [cf[:-1], cg], :668-669)
        const double v = src ? va : vb;
        bool tie;
        double dx;
        if (STRICT) {
            tie = !src && ia > 0 && (v == tprev);
            double xs = xa;
            if (tie) xs = lds64(xfa + 8u * (unsigned)(ia - 1));  // rare: predicated load
            dx = xs - xb;                                        // :671-672,676-677
        } else {
            tie = !src && (v == runf_val);        // target knot equal to the last consumed source knot
            const int eqf = (v == runf_val) ? runf_len : 0;
            const int eqg = (v == rung_val) ? rung_len : 0;
            const int indf = ia - eqf;                           // bisect_left(cf, v) (:671)
            const int indg = src ? ib : ib - eqg;                // bisect_left(cg, v) (:672)
            runf_len = src ? eqf + 1 : runf_len;
            rung_len = src ? rung_len : eqg + 1;
            rung_val = src ? rung_val : v;
            runf_val = src ? v : runf_val;
            dx = lds64(xfa + 8u * (unsigned)indf) - lds64(xfa + 8u * (unsigned)(npad + indg));   // :676-677
        }
        // next head of the consumed side (index into the contiguous [cf | cg] / [x_f | x_g] arrays)
        const int nxt = src ? ia + 1 : npad + ib + 1;
        const bool inr = src ? (ia + 1 < ia1) : (ib + 1 < ib1);
        double nv = CUDART_INF;
        if (inr) nv = lds64(cfa + 8u * (unsigned)nxt);
        double nx = 0.0;
        if (STRICT) nx = lds64(xfa + 8u * (unsigned)(src ? min(ia + 1, n - 1) : npad + min(ib + 1, m - 1)));
        acc.common += (tie && ib < m - 1) ? 1 : 0;               // np.intersect1d(cg[:-1], cf[:-1]) (:664)
        if (MO) { mo[k] = src ? ia : n - 1 + ib; ++k; }
        const double dt = v - tprev;                             // :673
        tprev = v;
        const double c1 = fabs(dx), c2 = dx * dx;
        if (PM & 1) {
            acc.w1 = fma(c1, dt, acc.w1);                        // :690
            // sign(dx) dt (:693): copy the sign of dx onto dt (dt >= 0), zero when dx == 0
            const int hi = __double2hiint(dt) ^ (__double2hiint(dx) & 0x80000000);
            const double sdt = __hiloint2double(hi, __double2loint(dt));
            acc.p1 += (dx != 0.0) ? sdt : 0.0;
        }
        if (PM & 2) {
            acc.w2 = fma(c2, dt, acc.w2);                        // :699-700
            acc.p2 = fma(dx, dt, acc.p2);                        // :703 (doubled once at the end: exact)
        }
        if (!STRICT && first) { first_c1 = c1; first_c2 = c2; first = false; }
        if (E1 || E2) {
            const double D1 = pc1 - c1, D2 = pc2 - c2;
            if (pj >= 0) {
                if (E1) sts64(e1a + 8u * (unsigned)pj, D1);
                if (E2) sts64(e2a + 8u * (unsigned)pj, D2);
            }
            if (E1) acc.z1 = fma(pcf, D1, acc.z1);
            if (E2) acc.z2 = fma(pcf, D2, acc.z2);
            pj = src ? ia : -1; pcf = src ? v : 0.0;             // pcf = 0 <=> nothing pending
            if (E1) pc1 = c1;
            if (E2) pc2 = c2;
        }
        if (STRICT) {
            xa = src ? nx : xa;
            xb = src ? xb : nx;
        }
        va = src ? nv : va;
        vb = src ? vb : nv;
        ia += src ? 1 : 0;
        ib += src ? 0 : 1;
    }

    // the chain's last knot, if a source knot, needs the first |dx|^p of the NEXT knot (0 past the end)
    __device__ __forceinline__ void finish(uint32_t e1a, uint32_t e2a, double nc1, double nc2, Acc& acc) {
        if ((E1 || E2) && pj >= 0) {
            const double D1 = pc1 - nc1, D2 = pc2 - nc2;
            if (E1) { sts64(e1a + 8u * (unsigned)pj, D1); acc.z1 = fma(pcf, D1, acc.z1); }
            if (E2) { sts64(e2a + 8u * (unsigned)pj, D2); acc.z2 = fma(pcf, D2, acc.z2); }
        }
    }
};

// merge-path split: how many source knots are among the first d merged knots (source first on ties)
__device__ __forceinline__ int merge_split(const double* cf, const double* cg, int n, int m, int d) {
    int lo = max(0, d - m), hi = min(d, n - 1);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cf[mid] <= cg[d - 1 - mid]) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// the splits of a lane's two chains side by side: two independent chains of dependent shared-memory loads per lane
// instead of one after the other (the bisection is pure load latency)
__device__ __forceinline__ void merge_split2(const double* cf, const double* cg, int n, int m, int dA, int dB,
                                             int& ia, int& ib) {
    int loA = max(0, dA - m), hiA = min(dA, n - 1);
    int loB = max(0, dB - m), hiB = min(dB, n - 1);
    while (loA < hiA || loB < hiB) {
        const int mA = (loA + hiA) >> 1, mB = (loB + hiB) >> 1;
        const bool actA = loA < hiA, actB = loB < hiB;
        const double fa = cf[mA], ga = cg[min(max(dA - 1 - mA, 0), m - 1)];    // in range whenever the search is active
        const double fb = cf[mB], gb = cg[min(max(dB - 1 - mB, 0), m - 1)];
        if (actA) { if (fa <= ga) loA = mA + 1; else hiA = mA; }
        if (actB) { if (fb <= gb) loB = mB + 1; else hiB = mB; }
    }
    ia = loA; ib = loB;
}

template <bool STRICT, int PM, int EM, bool MO>
__device__ __forceinline__ void warp_merge(const double* cf, const double* xf, int npad,
                                           double* e1s, double* e2s, int32_t* mo, int n, int m, int lane, Acc& acc) {
    const double* cg = cf + npad;
    const int K = n - 1 + m;
    const int per = (K + 63) >> 6;                // knots per chain (64 chains per warp)
    const int dA = min(2 * lane * per, K), dB = min(dA + per, K), dE = min(dB + per, K);
    int iaA, iaB;
    merge_split2(cf, cg, n, m, dA, dB, iaA, iaB);
    int iaE = __shfl_down_sync(kFull, iaA, 1);    // the next lane's first chain starts where this lane's second ends
    if (lane == 31) iaE = n - 1;
    constexpr bool E1 = (EM & 1) != 0, E2 = (EM & 2) != 0;
    Chain<STRICT, PM, EM, MO> A, B;
    A.init(cf, xf, npad, dA, iaA, iaB, dA - iaA, dB - iaB, n, m);
    B.init(cf, xf, npad, dB, iaB, iaE, dB - iaB, dE - iaE, n, m);
    __syncwarp();                                 // all look-back / split reads done before any lane parks an E_j
    const int lenA = dB - dA, lenB = dE - dB;     // lenB <= lenA
    const uint32_t cfa = (uint32_t)__cvta_generic_to_shared(cf), xfa = (uint32_t)__cvta_generic_to_shared(xf);
    const uint32_t e1a = E1 ? (uint32_t)__cvta_generic_to_shared(e1s) : 0u;
    const uint32_t e2a = E2 ? (uint32_t)__cvta_generic_to_shared(e2s) : 0u;
    int i = 0;
#pragma unroll 2
    for (; i < lenB; ++i) {
        A.step(cfa, xfa, npad, e1a, e2a, mo, n, m, acc);
        B.step(cfa, xfa, npad, e1a, e2a, mo, n, m, acc);
    }
    for (; i < lenA; ++i) A.step(cfa, xfa, npad, e1a, e2a, mo, n, m, acc);
    if (E1 || E2) {
        double nc1 = __shfl_down_sync(kFull, A.first_c1, 1), nc2 = __shfl_down_sync(kFull, A.first_c2, 1);
        if (lane == 31) { nc1 = 0.0; nc2 = 0.0; }
        A.finish(e1a, e2a, B.first_c1, B.first_c2, acc);      // an empty chain has first_c = 0 = |dx|^p past the end
        B.finish(e1a, e2a, nc1, nc2, acc);
    }
}

// dW_i = (sum_{j>=i} E_j - Z) / amp for one order: warp suffix scan over the parked E_j (slots >= n-1
// hold no E: cf[n-1] is not a knot), 16-byte coalesced stores.
__device__ __forceinline__ void suffix_phase(const double* es, double* out, double Z, double amp, int n, int npad,
                                             int lane) {
    const double ramp = 1.0 / amp;
    double carry = 0.0;
    for (int base = npad - 128; base >= 0; base -= 128) {
        const int idx = base + 4 * lane;
        const double2 q0 = *reinterpret_cast<const double2*>(es + idx);
        const double2 q1 = *reinterpret_cast<const double2*>(es + idx + 2);
        double e[4] = {q0.x, q0.y, q1.x, q1.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (idx + i >= n - 1) e[i] = 0.0;
        e[2] += e[3]; e[1] += e[2]; e[0] += e[1];
        double inc = e[0];                        // warp inclusive SUFFIX scan of the lane totals
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double o = __shfl_down_sync(kFull, inc, off);
            if (lane + off < 32) inc += o;
        }
        const double add = carry + (inc - e[0]);
        double r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = div_by((add + e[i]) - Z, amp, ramp);
        if ((idx + 3 < n) && ((n & 1) == 0)) {    // 16-byte aligned groups of 4
            *reinterpret_cast<double2*>(out + idx) = make_double2(r[0], r[1]);
            *reinterpret_cast<double2*>(out + idx + 2) = make_double2(r[2], r[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (idx + i < n) out[idx + i] = r[i];
        }
        carry += __shfl_sync(kFull, inc, 0);
    }
}

__global__ void __launch_bounds__(512, 1) k_ot1d_warp(Ot1dArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = a.n, m = a.m, npad = a.npad, mpad = a.mpad;
    double* const sx = reinterpret_cast<double*>(smem_raw);
    const size_t xbytes = a.xshared ? (size_t)(npad + mpad) * 8 : 0;
    double* const cf = reinterpret_cast<double*>(smem_raw + xbytes + (size_t)warp * a.per_warp);
    double* const cg = cf + npad;
    double* const e2buf = cg + mpad;              // present only if a.need_e2
    double* const xown = e2buf + (a.need_e2 ? npad : 0);
    const double* const xf = a.xshared ? sx : xown;
    if (a.xshared) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) sx[i] = a.xf[i];
        for (int i = threadIdx.x; i < m; i += blockDim.x) sx[npad + i] = a.xg[i];
        __syncthreads();
    }
    const int K = n - 1 + m;
    const bool want1 = a.deriv && a.dW1 && (a.pmask & 1), want2 = a.deriv && a.dW2 && (a.pmask & 2);
    int st_neg = 0, st_common = 0;

    for (long long b = (long long)blockIdx.x * a.wpc + warp; b < a.B; b += (long long)gridDim.x * a.wpc) {
        if (!a.xshared) {
            const double* gxf = a.xf + b * a.xf_stride;
            const double* gxg = a.xg + b * a.xg_stride;
            for (int i = lane; i < n; i += 32) xown[i] = gxf[i];
            for (int i = lane; i < m; i += 32) xown[npad + i] = gxg[i];
        }
        const int esz = a.dtype == WFOT_F32 ? 4 : 8;
        const void* frow = reinterpret_cast<const unsigned char*>(a.f) + b * a.f_stride * esz;
        const void* grow = reinterpret_cast<const unsigned char*>(a.g) + b * a.g_stride * esz;
        int negf = 0, negg = 0;
        bool strictf = false, strictg = false;
        const double amp = warp_cdf(frow, a.dtype, n, npad, cf, a.cdf_f ? a.cdf_f + b * n : nullptr, lane, negf, strictf);
        warp_cdf(grow, a.dtype, m, mpad, cg, a.cdf_g ? a.cdf_g + b * m : nullptr, lane, negg, strictg);
        st_neg += (__any_sync(kFull, negf) ? 1 : 0) + (__any_sync(kFull, negg) ? 1 : 0);
        __syncwarp();

        Acc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0};
        int32_t* const mo = a.merge_order ? a.merge_order + b * K : nullptr;
        const bool strict = strictf && strictg;
        // Specialisations: accumulated orders PM, derivative orders EM, merge-order output.  With the
        // merge order requested (a parity / debugging output) the general PM = 3 variant runs.
        int pm = a.pmask, em = (want1 ? 1 : 0) | (want2 ? 2 : 0);
        if (mo) { pm = 3; em = em ? 3 : 0; }
        // E^{(1)} (or the only order) overwrites cf in place; with both orders E^{(2)} goes to e2buf
        double* const e1s = cf;
        double* const e2s = (em == 3) ? e2buf : cf;
#define WFOT_MERGE(S, P, E, M) warp_merge<S, P, E, M>(cf, xf, npad, e1s, e2s, mo, n, m, lane, acc)
#define WFOT_MERGE_S(P, E, M) do { if (strict) WFOT_MERGE(true, P, E, M); else WFOT_MERGE(false, P, E, M); } while (0)
        if (mo) { if (em) WFOT_MERGE_S(3, 3, true); else WFOT_MERGE_S(3, 0, true); }
        else if (pm == 1) { if (em) WFOT_MERGE_S(1, 1, false); else WFOT_MERGE_S(1, 0, false); }
        else if (pm == 2) { if (em) WFOT_MERGE_S(2, 2, false); else WFOT_MERGE_S(2, 0, false); }
        else if (em == 3) WFOT_MERGE_S(3, 3, false);
        else if (em == 2) WFOT_MERGE_S(3, 2, false);
        else if (em == 1) WFOT_MERGE_S(3, 1, false);
        else WFOT_MERGE_S(3, 0, false);
#undef WFOT_MERGE_S
#undef WFOT_MERGE
        st_common += acc.common;
        const double w1 = warp_sum(acc.w1), w2 = warp_sum(acc.w2), p1 = warp_sum(acc.p1), p2 = 2.0 * warp_sum(acc.p2);
        if (lane == 0) {
            if (a.W) { if (a.pmask & 1) a.W[2 * b] = w1; if (a.pmask & 2) a.W[2 * b + 1] = w2; }
            if (a.dpos) { if (a.pmask & 1) a.dpos[2 * b] = p1; if (a.pmask & 2) a.dpos[2 * b + 1] = p2; }
            if (a.amp_f) a.amp_f[b] = amp;
        }

        // ---- dW_i = (sum_{j>=i} E_j - sum_j cf_j E_j) / amp   (:682-686,694,704 in O(n) form)
        if (em) {
            __syncwarp();                         // E_j parked by other lanes
            if (want1) suffix_phase(e1s, a.dW1 + b * n, warp_sum(acc.z1), amp, n, npad, lane);
            if (want2) suffix_phase(e2s, a.dW2 + b * n, warp_sum(acc.z2), amp, n, npad, lane);
        }
        __syncwarp();                             // cf / cg are rewritten by the next pair
    }
    if (a.status) {
        if (lane == 0 && st_neg) atomicAdd(a.status + WFOT_STAT_NEG_PDF, st_neg);
        const int c = __reduce_add_sync(kFull, st_common);
        if (lane == 0 && c) atomicAdd(a.status + WFOT_STAT_COMMON_CDF, c);
    }
}

}  // namespace wfot

using namespace wfot;

extern "C" int wfot_ot1d_batch(const void* f, const void* g, int in_dtype, const double* xf, const double* xg,
                               long long f_stride, long long g_stride, long long xf_stride, long long xg_stride,
                               int n, int m, int B, int pmask, int derivatives, double* W, double* dW1,
                               double* dW2, double* dpos, double* amp_f, double* cdf_f, double* cdf_g,
                               int32_t* merge_order, int32_t* status, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!f || !g || !xf || !xg || n < 1 || m < 1 || B <= 0 || pmask < 1 || pmask > 3 ||
        (in_dtype != WFOT_F32 && in_dtype != WFOT_F64))
        return WFOT_ERR_INVALID_ARG;
    Ot1dArgs a;
    a.f = f; a.g = g; a.dtype = in_dtype; a.xf = xf; a.xg = xg;
    a.f_stride = f_stride; a.g_stride = g_stride; a.xf_stride = xf_stride; a.xg_stride = xg_stride;
    a.n = n; a.m = m; a.pmask = pmask; a.deriv = derivatives; a.B = B;
    a.W = W; a.dW1 = dW1; a.dW2 = dW2; a.dpos = dpos; a.amp_f = amp_f; a.cdf_f = cdf_f; a.cdf_g = cdf_g;
    a.merge_order = merge_order; a.status = status;
    a.npad = ((n + 127) / 128) * 128;
    a.mpad = ((m + 127) / 128) * 128;
    a.xshared = (xf_stride == 0 && xg_stride == 0) ? 1 : 0;
    a.need_e2 = (derivatives && ((dW1 && dW2 && pmask == 3) || (merge_order && (dW1 || dW2)))) ? 1 : 0;
    const size_t xbytes = a.xshared ? (size_t)(a.npad + a.mpad) * 8 : 0;
    const size_t per_warp = (size_t)(a.npad + a.mpad) * 8 * (a.xshared ? 1 : 2) + (a.need_e2 ? (size_t)a.npad * 8 : 0);
    a.per_warp = per_warp;
    const size_t budget = 224 * 1024;               // per SM
    if (xbytes + per_warp > 220 * 1024) return WFOT_ERR_UNSUPPORTED;
    // one persistent CTA per SM with as many warps (= pairs in flight) as shared memory allows, <= 16
    int best_wpc = (int)((budget - 1024 - xbytes) / per_warp);
    if (best_wpc > 16) best_wpc = 16;
    if (best_wpc < 1) return WFOT_ERR_UNSUPPORTED;
    a.wpc = best_wpc;
    const size_t smem = xbytes + per_warp * (size_t)a.wpc;
    cudaError_t e = cudaFuncSetAttribute(k_ot1d_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_ot1d_warp)");
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ot1d_warp, 32 * a.wpc, smem);
    if (e != cudaSuccess || per_sm < 1) return cuda_fail(e, "k_ot1d_warp occupancy");
    int sms = wfot_device_sm_count();
    if (sms <= 0) sms = 148;
    long long grid = (long long)sms * per_sm;
    const long long need = ((long long)B + a.wpc - 1) / a.wpc;
    if (grid > need) grid = need;
    k_ot1d_warp<<<(int)grid, 32 * a.wpc, smem, stream>>>(a);
    note_launches(1);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "wfot_ot1d_batch launch");
    return WFOT_OK;
}
