// wfot_dev_options.h -- process-wide tuning switches behind include/wfot_dev.h (development and
// benchmarking only; every option defaults to 0 = the library's own choice).
#pragma once

namespace wfot {

enum DevOption {
    kOptPipeline = 0,      // fused path: 0 auto, 1 single-kernel form, 2 two-kernel (scan + resolve) form
    kOptResolveShape = 1,  // k_resolve: 0 auto, 1 = 128 registers x 2 CTAs/SM, 2 = 80 registers x 3 CTAs/SM, 5 = 128 threads, 6 CTAs/SM
    kOptFusedThreads = 2,  // k_misfit_grad threads per CTA: 0 auto, 64 / 128 / 256
    kOptClusterMax = 3,    // largest thread-block cluster per window: 0 auto (8), 1 = no clusters
    kOptTile = 4,          // argmin tile: 0 auto, 8 or 16 segments
    kOptSplitChunk = 5,    // windows per scan/resolve launch pair: 0 auto
    kOptOverlap = 6,       // two-kernel form: 0 auto (resolve of chunk c next to the scan of chunk c + 1), 1 sequential
    kOptScanShape = 7,     // k_scan: 0 auto, 2 = 256 threads x 2 per SM (128 registers), 3 = 256 x 3 (80 registers), 4 = 128 threads x 6
    kOptSkipKernel = 8,    // two-kernel form, timing aid: 1 = do not launch k_resolve, 2 = do not launch k_scan (stale scan results)
    kOptCount = 12
};

int dev_option(int id);

}  // namespace wfot
