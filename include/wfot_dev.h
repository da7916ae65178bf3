/*
 * wfot_dev.h -- development / measurement entry points of libwfot.so.  NOT part of the drop-in
 * boundary (include/wfot.h): nothing here is needed to run the hot path, and the Python shim
 * never calls it.  bench.py uses the FP32 probe as the roofline denominator of the CUDA-core
 * bound scan; scripts/ use the option switch for A/B timing of kernel variants.
 */
#ifndef WFOT_DEV_H
#define WFOT_DEV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Runs `iters` dependent-free FFMA2 (packed) or FFMA (scalar) bundles on every SM; returns the
 * executed FMA lane-operations through *fma_ops (host pointer).  FP32-pipe peak measurement. */
int wfot_fp32_peak_probe(int packed, int iters, float* sink, double* fma_ops, void* stream);

/* Process-wide tuning switch (see csrc/wfot_dev_options.h for the ids); value 0 restores the
 * library's own choice.  Returns the previous value, or -1 for an unknown id. */
int wfot_dev_set_option(int id, int value);

/* The two elementary functions of the fused path's density epilogue on arbitrary arguments (n device doubles):
 * exp_neg_out[i] = exp(-x[i]) (table + degree-5 polynomial, csrc/wfot_exp.cuh), rsqrt_out[i] = 1/sqrt(x[i])
 * (one MUFU + third-order correction).  For the accuracy tests. */
int wfot_dev_epilogue_math(const double* x, double* exp_neg_out, double* rsqrt_out, int n, void* stream);

/* Number of kernels the library has launched in this process so far (bench.py's gpu_launches). */
long long wfot_dev_kernel_launches(void);

/* While `iray` is non-NULL, every wfot_misfit_grad_batch / wfot_marginal_cdfs_batch call also writes the
 * nearest-segment index of every pixel of every window to iray (B, nug * ntg) int32 (device memory owned by the
 * caller, large enough for the largest call made).  bench.py uses it to check the TIMED kernels' indices
 * against the CPU oracle after the timed region.  NULL switches the capture off. */
void wfot_dev_capture_iray(int32_t* iray);

/* While `cycles` is non-NULL, k_resolve (two-kernel form) adds the SM clock cycles its CTAs spend per phase to
 * cycles[0..5] (uint64, device memory): [0] window preparation, [1] P1 per-pixel resolve + density, [2] P2 + P3
 * marginals and 1-D OT, [3] P4 gradient assembly, [4] windows, [5] unused.  Thread 0 of every CTA reads the clock
 * at the phase boundaries; for scripts/phase_cycles.py.  NULL switches it off. */
void wfot_dev_phase_cycles(unsigned long long* cycles);

#ifdef __cplusplus
}
#endif
#endif /* WFOT_DEV_H */
