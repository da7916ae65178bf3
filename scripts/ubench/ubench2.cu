// Register-bank experiments for FFMA2 operand forms (sm_100a).  Scalars come from an LDS.128
// destination quad so their register parity is known: q.x even, q.y odd, q.z even, q.w odd.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t f2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(512) k(int iters, float* out, long long* cyc, const float4* tab) {
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = tab[threadIdx.x];
    __syncthreads();
    uint64_t y[8], acc[8];
    for (int i = 0; i < 8; ++i) { y[i] = pk(1.f + i * 1e-3f + threadIdx.x * 1e-6f, 0.5f + i * 1e-3f); acc[i] = pk(0.f, 0.f); }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4 q = sm[(it * 4 + u) & 63];   // x:even y:odd z:even w:odd
            if (MODE == 0) {        // balanced: pair + (odd scalar, even scalar)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = f2(acc[i], pk(q.y, q.y), pk(q.z, q.z));
            } else if (MODE == 1) { // unbalanced: pair + (odd, odd)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = f2(acc[i], pk(q.y, q.y), pk(q.w, q.w));
            } else if (MODE == 2) { // unbalanced: pair + (even, even)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = f2(acc[i], pk(q.x, q.x), pk(q.z, q.z));
            } else if (MODE == 3) { // a*a + c with distinct pairs
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = f2(y[i], y[i], acc[(i + 3) & 7]);
            } else if (MODE == 4) { // a*a + c accumulate in place
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = f2(y[i], y[i], acc[i]);
            } else if (MODE == 5) { // pair*scalar + pair
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = f2(y[i], pk(q.y, q.y), acc[i]);
            } else if (MODE == 6) { // alternating balanced/unbalanced like the current loop
#pragma unroll
                for (int i = 0; i < 8; i += 2) { acc[i] = f2(acc[i], pk(q.y, q.y), pk(q.z, q.z)); acc[i + 1] = f2(acc[i + 1], pk(q.x, q.x), pk(q.w, q.w)); }
            } else if (MODE == 7) { // same scalar pair for all 8 (reuse-friendly), balanced
#pragma unroll
                for (int i = 0; i < 8; ++i) { acc[i] = f2(acc[i], pk(q.w, q.w), pk(q.x, q.x)); y[i] = f2(y[i], pk(q.y, q.y), pk(q.y, q.y)); }
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) { float a, b; upk(acc[i], a, b); s += a + b; }
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
int run(const char* name, int threads) {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc; float4* tab;
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cyc, sizeof(long long) * sms)); CK(cudaMalloc(&tab, 64 * 16));
    float h[256]; for (int i = 0; i < 256; ++i) h[i] = 1.0f + 1e-4f * i;
    CK(cudaMemcpy(tab, h, sizeof(h), cudaMemcpyHostToDevice));
    const int iters = 20000;
    k<MODE><<<sms, threads>>>(100, out, cyc, tab);
    CK(cudaDeviceSynchronize());
    k<MODE><<<sms, threads>>>(iters, out, cyc, tab);
    CK(cudaDeviceSynchronize());
    long long hc[256]; CK(cudaMemcpy(hc, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double c = 0; for (int i = 0; i < sms; ++i) c += hc[i]; c /= sms;
    const double winst = (double)iters * 4 * 8 * (threads / 128.0);
    printf("%-44s thr=%4d  clk per FFMA2 %.3f\n", name, threads, c / winst);
    cudaFree(out); cudaFree(cyc); cudaFree(tab);
    return 0;
}

int main() {
    for (int t : {256, 512}) {
        run<0>("pair + scalar(odd) + scalar(even)  balanced", t);
        run<1>("pair + scalar(odd) + scalar(odd)", t);
        run<2>("pair + scalar(even) + scalar(even)", t);
        run<3>("a*a + c (distinct pairs)", t);
        run<4>("a*a + acc (in place)", t);
        run<5>("pair*scalar + acc", t);
        run<6>("alternating (odd,even) / (even,odd) scalars", t);
        run<7>("mix: balanced + (odd,odd same reg) x2 per i", t);
    }
    return 0;
}
