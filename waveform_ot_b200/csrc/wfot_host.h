// wfot_host.h -- host-side helpers shared by the translation units of libwfot.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#include "wfot_device.cuh"
#include "wfot_dev_options.h"

namespace wfot {

constexpr int kFpChunk = 2048;   // windows prepared + scanned per launch pair (materialising path)

// Per-chunk scratch of the materialising fingerprint path (device memory).
struct FpWorkspace {
    WinHdr* hdr;     // [chunk]
    double2* pn;     // [chunk][nt]
    float4* A;       // [chunk][Spad]  {ex, ey, -am, -bm}
    float* H;        // [chunk][Spad]
    float4* bbox;    // [chunk][Spad / kTileMin]
    float* pxs;      // [chunk][ntg_pad]
    float* pys;      // [chunk][nug_pad]
    int Spad, ntg_pad, nug_pad;
};

inline int seg_pad(int nt) { return ((nt - 1 + kTilePad - 1) / kTilePad) * kTilePad; }
// argmin tile size: 8-segment tiles (fewer FP32 re-evaluations per pixel, finer pruning) unless the window is so long
// that the per-tile bookkeeping of the best-first walk would dominate
inline int tile_for(int nt) {
    if (const int o = dev_option(kOptTile)) return o == 8 ? 8 : 16;   // development override (wfot_dev.h)
    return (nt - 1 <= 2048) ? 8 : 16;
}
inline int pad4(int n) { return (n + 3) & ~3; }

inline size_t fp_workspace_per_window(int nt, int nug, int ntg) {
    const size_t Spad = (size_t)seg_pad(nt);
    return 128 + (size_t)nt * 16 + Spad * 20 + (Spad / kTileMin) * 16 + (size_t)(pad4(ntg) + pad4(nug)) * 4;
}

inline FpWorkspace fp_workspace_carve(void* base, int chunk, int nt, int nug, int ntg) {
    FpWorkspace ws;
    ws.Spad = seg_pad(nt);
    ws.ntg_pad = pad4(ntg);
    ws.nug_pad = pad4(nug);
    unsigned char* p = (unsigned char*)base;
    ws.hdr = (WinHdr*)p;   p += (size_t)chunk * 128;
    ws.pn = (double2*)p;   p += (size_t)chunk * nt * 16;
    ws.A = (float4*)p;     p += (size_t)chunk * ws.Spad * 16;
    ws.H = (float*)p;      p += (size_t)chunk * ws.Spad * 4;
    ws.bbox = (float4*)p;  p += (size_t)chunk * (ws.Spad / kTileMin) * 16;
    ws.pxs = (float*)p;    p += (size_t)chunk * ws.ntg_pad * 4;
    ws.pys = (float*)p;
    return ws;
}

extern thread_local char g_cuda_err[512];
int cuda_fail(cudaError_t e, const char* what);
// process-wide count of kernels launched by the library (wfot_dev_kernel_launches, bench.py's gpu_launches)
void note_launches(int n);

}  // namespace wfot
