"""ctypes binding of libwfot.so (include/wfot.h).  No CPU fallback: importing this
module fails loudly if the CUDA library has not been built
(`python -m waveform_ot_b200.build`)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WFOT_LIB_PATH") or os.path.join(HERE, "libwfot.so")     # override: A/B of two builds (scripts/)

STAT_NEG_PDF, STAT_COMMON_CDF, STAT_ZERO_DIST, STAT_DEGENERATE_SEG, STAT_SLOW_PIXELS = range(5)
STAT_SCAN_TILES = 6          # slots 6-7: one 64-bit counter
STAT_SLOTS = 8
F32, F64 = 0, 1
ERR_INVALID_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_WORKSPACE = -1, -2, -3, -4
W1, W2, W12 = 1, 2, 3


class wfot_grid(C.Structure):
    _fields_ = [("t0", C.c_double), ("t1", C.c_double), ("u0", C.c_double), ("u1", C.c_double),
                ("fp_t0", C.c_double), ("fp_t1", C.c_double), ("fp_u0", C.c_double), ("fp_u1", C.c_double),
                ("tantheta", C.c_double), ("has_fpgrid", C.c_int32), ("reserved", C.c_int32)]


GRID_DOUBLES = C.sizeof(wfot_grid) // 8   # the struct is 10 x 8 bytes; packed as float64 rows on device

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "waveform_ot_b200: %s is missing; build it with `python -m waveform_ot_b200.build` "
        "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_d = C.c_double
_sz = C.c_size_t

SIGNATURES = {
    "wfot_version": (C.c_int, []),
    "wfot_strerror": (C.c_char_p, [_i]),
    "wfot_last_cuda_error": (C.c_char_p, []),
    "wfot_device_sm_count": (C.c_int, []),
    "wfot_device_cc": (C.c_int, []),
    "wfot_fingerprint_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "wfot_fingerprint_batch": (C.c_int, [_p, _p, _i, _ll, _i, _p, _i, _i, _i, _i, _d, _i,
                                         _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p, _p]),
    "wfot_marginals_batch": (C.c_int, [_p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "wfot_otpdf1d_batch": (C.c_int, [_p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "wfot_ot1d_batch": (C.c_int, [_p, _p, _i, _p, _p, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _i,
                                  _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "wfot_plan_batch": (C.c_int, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "wfot_pdfderiv_batch": (C.c_int, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _d, _i, _p, _p]),
    "wfot_misfit_grad_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "wfot_misfit_grad_batch": (C.c_int, [_p, _p, _i, _ll, _i, _p, _i, _i, _i, _i, _d, _i, _i, _i,
                                         _p, _p, _p, _p, _i, _p, _p, _p, _p, _sz, _p, _p]),
    "wfot_marginal_cdfs_batch": (C.c_int, [_p, _p, _i, _ll, _i, _p, _i, _i, _i, _i, _d, _i, _i,
                                           _p, _p, _p, _p, _sz, _p, _p]),
    "wfot_ricker_batch": (C.c_int, [_p, _i, _d, _d, _p, _p, _p, _p]),
    "wfot_chain_batch": (C.c_int, [_p, _p, _i, _i, _i, _ll, _p, _p]),
    "wfot_sum_windows_workspace_bytes": (_sz, [_i]),
    "wfot_sum_windows": (C.c_int, [_p, _ll, _i, _p, _p, _sz, _p]),
}

# include/wfot_dev.h: measurement / tuning entry points, not part of the drop-in boundary
DEV_SIGNATURES = {
    "wfot_fp32_peak_probe": (C.c_int, [_i, _i, _p, _p, _p]),
    "wfot_dev_set_option": (C.c_int, [_i, _i]),
    "wfot_dev_capture_iray": (None, [_p]),
    "wfot_dev_phase_cycles": (None, [_p]),
    "wfot_dev_kernel_launches": (C.c_longlong, []),
    "wfot_dev_epilogue_math": (C.c_int, [_p, _p, _p, _i, _p]),
}
(OPT_PIPELINE, OPT_RESOLVE_SHAPE, OPT_FUSED_THREADS, OPT_CLUSTER_MAX, OPT_TILE, OPT_SPLIT_CHUNK, OPT_OVERLAP,
 OPT_SCAN_SHAPE, OPT_SKIP_KERNEL) = range(9)

for _name, (_res, _args) in list(SIGNATURES.items()) + list(DEV_SIGNATURES.items()):
    _f = getattr(lib, _name)       # AttributeError here = header/library mismatch
    _f.restype = _res
    _f.argtypes = _args


class WfotError(RuntimeError):
    pass


def check(status, what=""):
    if status != 0:
        msg = lib.wfot_strerror(status).decode()
        cuda = lib.wfot_last_cuda_error().decode()
        raise WfotError("%s failed: %s%s" % (what or "libwfot call", msg, (" [" + cuda + "]") if cuda else ""))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
