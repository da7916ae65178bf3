"""CPU: libwfot.so builds (nvcc cross-compiles sm_100a without a GPU), loads, and exports every
function include/wfot.h declares.  No compute calls."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header="wfot.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wfot_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header():
    from waveform_ot_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    names = _declared() + _declared("wfot_dev.h")
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.wfot_version.restype = ctypes.c_int
    assert lib.wfot_version() == 100
    lib.wfot_strerror.restype = ctypes.c_char_p
    assert lib.wfot_strerror(-4) == b"workspace too small"


def test_binding_table_matches_header():
    from waveform_ot_b200 import _cabi
    assert sorted(_cabi.SIGNATURES) == _declared()
    assert sorted(_cabi.DEV_SIGNATURES) == _declared("wfot_dev.h")
    # the drop-in boundary holds no probes or tuning switches
    assert not [n for n in _declared() if "probe" in n or "dev_" in n]
    assert ctypes.sizeof(_cabi.wfot_grid) == 80


def test_sass_is_blackwell_packed_fp32():
    """The scan loop must be the packed-FP32 sm_100a code path (FFMA2/FMUL2/FMNMX3), not a generic build."""
    import shutil
    import subprocess
    from waveform_ot_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    for op in ("FFMA2", "FMUL2", "FMNMX3", "FADD.SAT",
               "VOTE.ANY",              # per-lane tile pruning of the scan: one vote per tile
               "CREDUX.MIN",            # best-first tile order: one warp minimum pops the next tile
               "REDG.E.ADD.F64",       # gradient assembly: native FP64 reductions in L2
               "UCGABAR_ARV"):          # thread-block cluster barrier (one window per cluster for small batches)
        assert op in out, op
