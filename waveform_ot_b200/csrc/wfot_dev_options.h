// wfot_dev_options.h -- process-wide tuning switches behind include/wfot_dev.h (development and
// benchmarking only; every option defaults to 0 = the library's own choice).
#pragma once

namespace wfot {

enum DevOption {
    kOptPipeline = 0,      // fused path: 0 auto, 1 single-kernel form, 2 two-kernel (scan + resolve) form
    kOptResolveShape = 1,  // k_resolve CTA shape: 0 auto, 1 = 256 thr x 2/SM, 2 = 256 x 3, 3 = 512 x 2
    kOptFusedThreads = 2,  // k_misfit_grad threads per CTA: 0 auto, 64 / 128 / 256
    kOptClusterMax = 3,    // largest thread-block cluster per window: 0 auto (8), 1 = no clusters
    kOptTile = 4,          // argmin tile: 0 auto, 8 or 16 segments
    kOptSplitChunk = 5,    // windows per scan/resolve launch pair: 0 auto
    kOptCount = 8
};

int dev_option(int id);

}  // namespace wfot
