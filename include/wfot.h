/*
 * wfot.h -- C ABI of libwfot.so: the B200 (sm_100a) implementation of
 * waveform-ot's fingerprint + marginal-Wasserstein misfit hot path.
 *
 * The reference (msambridge/waveform-ot) has no FFI: its boundary is the
 * Python module surface libs/FingerprintLib.py + libs/OTlib.py.  Each entry
 * point below names the reference function(s) it replaces; the Python shim
 * (waveform_ot_b200/FingerprintLib.py, OTlib.py) binds them with ctypes and
 * re-creates the reference's classes/attributes on top.  INTEGRATION.md shows
 * the binding a reference maintainer would add.
 *
 * Conventions
 *  - plain C types only; every array pointer is a DEVICE pointer unless the
 *    function name ends in _host; `stream` is a cudaStream_t passed as void*.
 *  - the caller owns every buffer, including the scratch workspace whose size
 *    the matching *_workspace_bytes() call reports.  The library keeps no
 *    global state and never allocates persistent device memory.
 *  - calls are asynchronous on `stream`; return value 0 = launched OK,
 *    negative = wfot_status (see wfot_strerror).  Data-dependent conditions the
 *    reference reports as Python exceptions (negative density, common CDF
 *    values, zero distance in a derivative) are counted into the `status`
 *    device array (WFOT_STAT_* slots, int32 each, per call, accumulated with
 *    atomics) and mapped back to the reference's exception classes by the shim.
 *  - window b of a batch reads t + b*t_stride and w + b*nt (t_stride = 0
 *    shares one time axis); window b uses grids[b % n_grids] (1 = shared, B = one per
 *    window, nr*nc = one per station/component of every trial model).
 *  - pixel flat index k = iu*ntg + it (row-major (nug, ntg), the reference's
 *    meshgrid 'xy' order, libs/FingerprintLib.py:254-255).
 */
#ifndef WFOT_H
#define WFOT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WFOT_VERSION 100

/* ---- status codes ------------------------------------------------------ */
enum wfot_status {
    WFOT_OK = 0,
    WFOT_ERR_INVALID_ARG = -1,
    WFOT_ERR_CUDA = -2,
    WFOT_ERR_UNSUPPORTED = -3,  /* e.g. not an sm_100 device, window too large for smem */
    WFOT_ERR_WORKSPACE = -4
};

/* slots of the per-call `status` device array (int32[WFOT_STAT_SLOTS]) */
enum wfot_stat_slot {
    WFOT_STAT_NEG_PDF = 0,        /* OTpdf: min(pdf) < 0        (libs/OTlib.py:91)      */
    WFOT_STAT_COMMON_CDF = 1,     /* wasser: cf[:-1] n cg[:-1]  (libs/OTlib.py:663-666) */
    WFOT_STAT_ZERO_DIST = 2,      /* wdistderiv: d == 0 -> NaN  (libs/FingerprintLib.py:355) */
    WFOT_STAT_DEGENERATE_SEG = 3, /* zero-length segment        (libs/FingerprintLib.py:257, 0/0) */
    WFOT_STAT_SLOW_PIXELS = 4,    /* pixels resolved by the full FP64 rescan (diagnostic) */
    WFOT_STAT_SCAN_TILES = 6,     /* slots 6-7: 64-bit count of (warp, segment tile) pairs the pruned scan
                                     evaluated, in units of 2048 (pixel, segment) pairs (diagnostic) */
    WFOT_STAT_SLOTS = 8
};

enum wfot_dtype { WFOT_F32 = 0, WFOT_F64 = 1 };

/* p-mask for the Wasserstein order: the reference's distfunc strings */
enum wfot_pmask { WFOT_W1 = 1, WFOT_W2 = 2, WFOT_W12 = 3 };

/* The reference's `grid` tuple (t0,t1,u0,u1,Nu,Nt) plus optional fpgrid and
 * tan(theta): libs/FingerprintLib.py:53,75-106.  Nu/Nt are per call. */
typedef struct wfot_grid {
    double t0, t1, u0, u1;             /* non-dimensionalisation box            */
    double fp_t0, fp_t1, fp_u0, fp_u1; /* fingerprint box, used iff has_fpgrid  */
    double tantheta;                   /* metric weighting, 1.0 for theta = 45  */
    int32_t has_fpgrid;
    int32_t reserved;
} wfot_grid;

/* ---- library ------------------------------------------------------------ */
int wfot_version(void);
const char* wfot_strerror(int status);
/* text of the last CUDA error seen by the calling thread ("" if none) */
const char* wfot_last_cuda_error(void);
/* SM count / compute capability of the current device; < 0 on error */
int wfot_device_sm_count(void);
int wfot_device_cc(void);

/* ---- fingerprint (materialising path) ------------------------------------
 * Replaces waveformFP.__init__ + calcpdf(method='Enumerate') + wdist +
 * wdistderiv: libs/FingerprintLib.py:53-115, 117-177, 230-269, 333-385.
 * FP32 argmin over the segments with exact pruning of segment tiles that cannot
 * hold the minimum, then exact FP64 re-evaluation (in the reference's operation
 * order) of every near-minimal candidate, so `iray` equals the reference's
 * np.argmin first-minimum index and dfield / lray / xray are bit-identical.
 * Any output pointer may be NULL (not materialised).  Outputs are FP64:
 *   pn     (B, nt, 2)   normalised sample coordinates            (:110)
 *   dfield (B, nug, ntg) nearest distance                        (:265)
 *   iray   (B, nug*ntg) int32 nearest segment                    (:266)
 *   lray   (B, nug*ntg) clipped segment parameter                (:268)
 *   xray   (B, nug*ntg, 2) nearest point on the waveform         (:267)
 *   pdf    (B, nug, ntg) exp(-|d|/lambda) (q=0) or exp(-d^2/lambda) (q=2)  (:174,176)
 *   dddy   (B, nug*ntg, 2) d(d)/d(raw amplitude of the segment's end samples) (:385)
 * Size limit: the FP32 segment table of a window stays in one SM's shared memory
 * (20 B per sample), i.e. nt up to about 10 000; beyond that WFOT_ERR_UNSUPPORTED.
 */
size_t wfot_fingerprint_workspace_bytes(int B, int nt, int nug, int ntg);
int wfot_fingerprint_batch(const void* t, const void* w, int in_dtype, long long t_stride, int nt,
                           const wfot_grid* grids, int n_grids, int B, int nug, int ntg,
                           double lambda, int q,
                           double* pn, double* dfield, int32_t* iray, double* lray, double* xray,
                           double* pdf, double* dddy,
                           void* workspace, size_t workspace_bytes, int32_t* status, void* stream);

/* ---- marginals ------------------------------------------------------------
 * Replaces OTpdf.__init__ (2-D) + setMarginals: libs/OTlib.py:90-117, 146-163.
 * pdf (B, nug, ntg) FP64 un-normalised ->
 *   amp (B,), marg_t (B, ntg) = sum over rows / amp, marg_u (B, nug) = sum over
 *   columns / amp  (i.e. the marginals of the normalised 2-D density).
 * Vectorised, coalesced, deterministic (fixed-order) reductions. */
int wfot_marginals_batch(const double* pdf, int B, int nug, int ntg,
                         double* amp, double* marg_t, double* marg_u,
                         int32_t* status, void* stream);

/* ---- 1-D OTpdf ---------------------------------------------------------------
 * Replaces OTpdf.__init__ for 1-D input: libs/OTlib.py:91-93,112-114.
 * f (B, n) un-normalised -> amp (B,), pdf_norm (B, n) = f/amp, cdf (B, n) =
 * cumsum(pdf_norm)/cumsum(pdf_norm)[-1] (FP64 block prefix scan).  Outputs nullable. */
int wfot_otpdf1d_batch(const void* f, int in_dtype, int n, int B,
                       double* amp, double* pdf_norm, double* cdf, int32_t* status, void* stream);

/* ---- 1-D optimal transport -------------------------------------------------
 * Replaces OTpdf.__init__ (1-D) + wasser(distfunc in {'W1','W2','W12'},
 * derivatives=...): libs/OTlib.py:90-117, 596-706.
 * f (B, n) un-normalised source amplitudes, g (B, m) target; *_stride in
 * elements (0 = shared by the whole batch).  Prefix-scan CDFs (FP64), stable
 * merge of cf[:-1] and cg (source first on ties), bisect_left quantile ranks.
 * Outputs (any may be NULL): W (B, 2) = [W1, W2^2] (slot unused by pmask left
 * untouched), dW1/dW2 (B, n) d/d(un-normalised f), dpos (B, 2) d/d(translation
 * of the source), amp_f (B,), cdf_f (B, n), cdf_g (B, m),
 * merge_order (B, n+m-1) int32 = the reference's `tkarg` (:669).
 * WFOT_STAT_COMMON_CDF counts exact cf/cg coincidences (:663-666). */
int wfot_ot1d_batch(const void* f, const void* g, int in_dtype,
                    const double* xf, const double* xg,
                    long long f_stride, long long g_stride, long long xf_stride, long long xg_stride,
                    int n, int m, int B, int pmask, int derivatives,
                    double* W, double* dW1, double* dW2, double* dpos,
                    double* amp_f, double* cdf_f, double* cdf_g, int32_t* merge_order,
                    int32_t* status, void* stream);

/* ---- transport plan -----------------------------------------------------------
 * Replaces the returnplan branch of wasser(): libs/OTlib.py:718-740, and the per-slice
 * scatter of SlicedWasserstein(returnplan / calcWplan): libs/OTlib.py:1247-1262.
 * From the outputs of wfot_ot1d_batch (cdf_f (B, n), cdf_g (B, m), merge_order (B, n+m-1),
 * amp_f (B,)):  H (n, m) per pair with H[indf_k, indg_k] += dt_k over the merged knots, and
 * (dH != NULL) dH (n, n, m) per pair with dH[l, indf_k, indg_k] += Diffdtk[l, k] (:682-686,
 * 731-733), indf / indg = bisect_left ranks of the knot in the two CDFs.
 * perm_f (B, n) / perm_g (B, m) (nullable int32): rows / columns (and the derivative index l)
 * are scattered through these permutations - the argsort of a slice's projected positions.
 * accumulate != 0: all B pairs add into ONE H (n, m) / dH (n, n, m) (the slice sum).
 * H and dH are zeroed by the call. */
int wfot_plan_batch(const double* cdf_f, const double* cdf_g, const int32_t* merge_order,
                    const double* amp_f, const int32_t* perm_f, const int32_t* perm_g,
                    int n, int m, int B, int accumulate, double* H, double* dH, void* stream);

/* ---- gradient assembly (materialising path) --------------------------------
 * Replaces waveformFP.PDFderiv / PDFderivMarg: libs/FingerprintLib.py:182-228.
 * out (B, nchain, nt) = -1/lambda * segmented sum over pixels keyed by iray of
 * dddy * pdf * chain (* 2|d| if q == 2).  chain (B, nchain, nug*ntg) or NULL
 * (chain == 1, nchain must be 1). */
int wfot_pdfderiv_batch(const double* pdf, const double* dfield, const int32_t* iray,
                        const double* dddy, const double* chain, int nchain,
                        int B, int npix, int nt, double lambda, int q,
                        double* out, void* stream);

/* ---- Ricker forward model ------------------------
 * Replaces ru.rickerwavelet(tpert, amp, f, trange=(t0,t1), deriv=...) for noise-free
 * waveforms: libs/ricker_util.py:22-30, 38-89.  params (M, 3) FP64 rows = (tpert, amp, f).
 * Outputs: t (M, 256) sample times, w (M, 256) amplitudes, dw (M, 3, 256) (NULL = no
 * derivatives) = d(w)/d(time offset, amplitude, frequency factor) (:84-86). */
int wfot_ricker_batch(const double* params, int M, double t0, double t1,
                      double* t, double* w, double* dw, void* stream);

/* ---- fused misfit + gradient (throughput path) -------------------------------
 * One "evaluation" per window: fingerprint -> marginals -> W_p^p per marginal ->
 * d/d(waveform amplitudes) per marginal and d/d(window position), nothing but
 * the waveform read from and the (2 + 2*nt + 1) results written to HBM.
 * Replaces the chain ru.BuildOTobjfromWaveform -> OT.MargWasserstein(derivatives
 * =True, returnmargW=True) -> wf.PDFderivMarg: libs/ricker_util.py:204-268,
 * 321-337; libs/OTlib.py:1055-1154; libs/FingerprintLib.py:205-228.
 * Target (observed) marginals are given as their CDFs + bin positions:
 *   tgt_cdf_t/tgt_x_t (tgt_rows, ntg), tgt_cdf_u/tgt_x_u (tgt_rows, nug), window b
 *   uses row b % tgt_rows (1 = one observation for the whole batch, B = one per
 *   window, nr*nc = one per station/component shared by all trial models).
 * Outputs: W (B, 2) = [W^t, W^u]; grad (B, 2, nt) = [dW^t/dw, dW^u/dw] (NULL =
 * misfit only); dwg (B,) = dW^t/d(translation) in normalised time units
 * (divide by tan(theta)*(t1-t0) as libs/ricker_util.py:333).
 * pmask is WFOT_W1 or WFOT_W2; misfit-only calls (grad == NULL) also take WFOT_W12: both orders
 * from one fingerprint (the misfit surfaces of Ricker_Figs_1_7.ipynb cells 34/38 want W1 and W2 of
 * the same windows), W then (B, 4) = [W1^t, W1^u, W2^t, W2^u] and dwg (B, 2) = [of W1^t, of W2^t].
 * If transform != 0 the arctan amplitude
 * transform of libs/ricker_util.py:270-275 is applied in-kernel with each
 * grid's (u0,u1) and the gradient is multiplied by d(un)/du (:393-397).
 * Reproducibility: W and dwg are bit-identical from run to run (fixed summation orders).  The
 * marginal CDFs - the quantities the reference compares for exact equality - are in addition
 * independent of the launch shape (single kernel, two kernels, clusters, threads per CTA); the
 * final sums of W over the merged knots are block reductions, so W / dwg agree across launch
 * shapes with different CTA sizes to ~1e-15 relative, not bitwise.  grad is assembled with
 * FP64 reductions in L2 in arrival order (one per run of pixels of a column that feed the same
 * sample): its last bits vary from run to run, by <= 1e-12 relative to the row's largest
 * entry (tests/test_gpu_parity.py::test_fused_run_to_run).  One of the two per-pixel weights
 * travels through the scratch slab with a 36-bit mantissa: grad agrees with the reference to
 * ~1e-11 relative to the row's largest entry (1e-13 measured), not to the last digit.
 * Size limit: sample coordinates, segment table, marginals and the OT scratch of a
 * window share one SM's shared memory (about 36 B per sample + 40 B per grid point
 * of the longer axis), i.e. nt up to about 6 000; beyond that WFOT_ERR_UNSUPPORTED
 * (use the materialising entry points above). */
size_t wfot_misfit_grad_workspace_bytes(int B, int nt, int nug, int ntg);
int wfot_misfit_grad_batch(const void* t, const void* w, int in_dtype, long long t_stride, int nt,
                           const wfot_grid* grids, int n_grids, int B, int nug, int ntg,
                           double lambda, int q, int pmask, int transform,
                           const double* tgt_cdf_t, const double* tgt_x_t,
                           const double* tgt_cdf_u, const double* tgt_x_u, int tgt_rows,
                           double* W, double* grad, double* dwg,
                           void* workspace, size_t workspace_bytes, int32_t* status, void* stream);

/* ---- observed window: CDFs of the two marginals --------------------------------
 * What the fused entry point above consumes as its target.  Replaces, for the OBSERVED
 * window(s), the chain ru.BuildOTobjfromWaveform (fingerprint, libs/ricker_util.py:204-268)
 * -> OTpdf (2-D, libs/OTlib.py:90-93) -> setMarginals (:146-160) -> OTpdf of each marginal
 * (:91-93,112-114): cdf_t (B, ntg), cdf_u (B, nug), amp (B,) = sum of the 2-D density
 * (nullable).  It runs the SAME kernels and the same summation orders as
 * wfot_misfit_grad_batch (which derives the predicted window's CDFs on the fly), and every
 * sum that feeds a CDF bit is independent of the launch shape, so a predicted window that
 * equals the observed one produces bit-identical CDFs and is reported through
 * WFOT_STAT_COMMON_CDF exactly as the reference raises TargetSourceCDFError (:663-666).
 * The bin positions the target also needs are the pixel axes (np.linspace of the window's
 * normalised limits, libs/FingerprintLib.py:254), host arithmetic.
 * Workspace: wfot_misfit_grad_workspace_bytes(B, nt, nug, ntg). */
int wfot_marginal_cdfs_batch(const void* t, const void* w, int in_dtype, long long t_stride, int nt,
                             const wfot_grid* grids, int n_grids, int B, int nug, int ntg,
                             double lambda, int q, int transform,
                             double* cdf_t, double* cdf_u, double* amp,
                             void* workspace, size_t workspace_bytes, int32_t* status, void* stream);

/* ---- chain to model parameters ----------------------------------------------
 * Replaces `dw.dot(dr)` / `d.dot(dr.flatten())`: libs/ricker_util.py:399-400,
 * libs/loc_cmt_util.py:283-296.  out (M, P) = J (M, P, L) . dr (M, L);
 * J_stride_models = 0 shares one Jacobian. */
int wfot_chain_batch(const double* J, const double* dr, int P, int L, int M,
                     long long J_stride_models, double* out, void* stream);

/* ---- batch reduction -----------------------------------------------------------
 * out (C,) = sum over b of in (B, C), FP64, fixed summation order: the same input gives the
 * same bits on every run and for every shard split that keeps the block structure (the
 * gradient rows fed to it carry their own ~1e-15 run-to-run noise, see above).  This is the local step before the single NCCL allreduce of
 * [sum misfit, sum gradient] across GPUs; the reference's counterpart is
 * the Python accumulation `mis += w2p` in libs/loc_cmt_util.py:260-271. */
size_t wfot_sum_windows_workspace_bytes(int C);
int wfot_sum_windows(const double* in, long long B, int C, double* out,
                     void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WFOT_H */
