// Instruction-throughput microbenchmarks for the scan loop's instruction mix (sm_100a).
// Each kernel runs an unrolled block of independent operations ITER times on every warp;
// throughput = executed warp-instructions / (elapsed SM cycles * 4 SMSPs).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

template <int MODE>
__global__ void __launch_bounds__(1024) k(int iters, float* out, long long* cyc) {
    // 12 accumulators (pairs), 4 read-only operand pairs
    uint64_t acc[12];
    float sc[8];
    for (int i = 0; i < 12; ++i) acc[i] = pk(1.0f + threadIdx.x * 1e-6f + i, 0.5f + i);
    for (int i = 0; i < 8; ++i) sc[i] = 1.0f + 1e-7f * (i + threadIdx.x);
    uint64_t o0 = pk(1.0000001f, 0.9999999f), o1 = pk(1e-8f, -1e-8f), o2 = pk(0.99f, 1.01f), o3 = pk(1e-6f, 2e-6f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0) {   // FFMA2: acc = acc*o0 + o1   (2 shared operand pairs -> reuse cache friendly)
#pragma unroll
                for (int i = 0; i < 12; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(o0), "l"(o1));
            } else if (MODE == 1) {   // FFMA2 with three distinct, rotating register pairs
#pragma unroll
                for (int i = 0; i < 12; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc[i]) : "l"(acc[(i + 5) % 12]), "l"(acc[(i + 7) % 12]), "l"(acc[(i + 3) % 12]));
            } else if (MODE == 2) {   // FFMA2 with scalar-broadcast first operand (as in the scan loop)
#pragma unroll
                for (int i = 0; i < 12; ++i) { uint64_t b = pk(sc[i & 7], sc[i & 7]); asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc[i]) : "l"(b), "l"(acc[(i + 7) % 12]), "l"(acc[(i + 3) % 12])); }
            } else if (MODE == 3) {   // FADD.SAT with |x|
#pragma unroll
                for (int i = 0; i < 12; ++i) { float a, b; upk(acc[i], a, b); float r0, r1;
                    asm volatile("{ .reg .f32 t; abs.f32 t, %1; sub.sat.f32 %0, t, %2; }" : "=f"(r0) : "f"(a), "f"(sc[i & 7]));
                    asm volatile("{ .reg .f32 t; abs.f32 t, %1; sub.sat.f32 %0, t, %2; }" : "=f"(r1) : "f"(b), "f"(sc[(i + 1) & 7]));
                    acc[i] = pk(r0, r1); }
            } else if (MODE == 4) {   // FMNMX3
#pragma unroll
                for (int i = 0; i < 12; ++i) { float a, b, c, d; upk(acc[i], a, b); upk(acc[(i + 5) % 12], c, d); float r0, r1;
                    asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r0) : "f"(a), "f"(c), "f"(sc[i & 7]));
                    asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r1) : "f"(b), "f"(d), "f"(sc[(i + 3) & 7]));
                    acc[i] = pk(r0, r1); }
            } else if (MODE == 5) {   // FMUL2 x*x
#pragma unroll
                for (int i = 0; i < 12; ++i) asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(acc[i]) : "l"(acc[(i + 5) % 12]));
            } else if (MODE == 6) {   // scalar FFMA 3 distinct regs
#pragma unroll
                for (int i = 0; i < 12; ++i) { float a, b, c, d; upk(acc[i], a, b); upk(acc[(i + 5) % 12], c, d); float r0, r1;
                    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r0) : "f"(a), "f"(c), "f"(d));
                    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r1) : "f"(b), "f"(d), "f"(c));
                    acc[i] = pk(r0, r1); }
            } else if (MODE == 7) {   // the scan loop's mix for 2 pixels x 6 "rows": 3 FFMA2 + FMUL2 + 2 FADD.SAT + 2 FMNMX(3)
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    uint64_t y2 = pk(sc[i], sc[i]), ny2 = pk(-sc[i], -sc[i]);
                    uint64_t al, pe, D;
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(al) : "l"(y2), "l"(o0), "l"(o2));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pe) : "l"(ny2), "l"(o1), "l"(o3));
                    float a, b; upk(al, a, b); float r0, r1;
                    asm volatile("{ .reg .f32 t; abs.f32 t, %1; sub.sat.f32 %0, t, %2; }" : "=f"(r0) : "f"(a), "f"(sc[7]));
                    asm volatile("{ .reg .f32 t; abs.f32 t, %1; sub.sat.f32 %0, t, %2; }" : "=f"(r1) : "f"(b), "f"(sc[7]));
                    uint64_t u2 = pk(r0, r1);
                    asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(D) : "l"(pe));
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(D) : "l"(u2), "l"(D));
                    float d0, d1, m0, m1; upk(D, d0, d1); upk(acc[i], m0, m1);
                    asm volatile("min.f32 %0, %0, %1;" : "+f"(m0) : "f"(d0));
                    asm volatile("min.f32 %0, %0, %1;" : "+f"(m1) : "f"(d1));
                    acc[i] = pk(m0, m1);
                    // perturb operands so nothing is loop-invariant
                    o2 = al;
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 12; ++i) { float a, b; upk(acc[i], a, b); s += a + b; }
    float a, b; upk(o2, a, b); s += a + b;
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
int run(const char* name, int warp_inst_per_iter, int threads, int ctas_per_sm) {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cyc, sizeof(long long) * sms * ctas_per_sm));
    const int iters = 20000;
    k<MODE><<<sms * ctas_per_sm, threads>>>(100, out, cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms * ctas_per_sm, threads>>>(iters, out, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[2048]; CK(cudaMemcpy(h, cyc, sizeof(long long) * sms * ctas_per_sm, cudaMemcpyDeviceToHost));
    double c = 0; for (int i = 0; i < sms * ctas_per_sm; ++i) c += h[i]; c /= sms * ctas_per_sm;
    const double warps_per_smsp = threads / 32.0 * ctas_per_sm / 4.0;
    const double winst = (double)iters * 4 * warp_inst_per_iter * warps_per_smsp;   // per SMSP
    printf("%-34s thr=%4d x%d  %.3f ms  cyc %.0f  warp-inst/clk/SMSP %.3f  (clk per warp-inst %.3f)  GHz %.3f\n", name, threads, ctas_per_sm, ms, c,
           winst / c, c / winst, c / (ms * 1e6));
    cudaFree(out); cudaFree(cyc);
    return 0;
}

int main() {
    for (int t : {128, 256, 512, 1024}) {
        printf("---- %d threads per SM (%d warps/SMSP)\n", t, t / 128);
        run<0>("FFMA2 acc*=o0+o1 (reuse)", 12, t, 1);
        run<1>("FFMA2 3 distinct pairs", 12, t, 1);
        run<2>("FFMA2 scalar-bcast + 2 pairs", 12, t, 1);
        run<3>("FADD.SAT |x|-h", 24, t, 1);
        run<4>("FMNMX3", 24, t, 1);
        run<5>("FMUL2 x*x", 12, t, 1);
        run<6>("FFMA scalar 3 regs", 24, t, 1);
        run<7>("scan mix (8 inst per 2 px)", 48, t, 1);
    }
    return 0;
}
