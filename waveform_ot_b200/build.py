"""Build libwfot.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwfot.so")
SOURCES = ["wfot_kernels.cu", "wfot_fused.cu", "wfot_split.cu", "wfot_ot1d.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(HERE, "..", "include", "wfot.h"), os.path.join(HERE, "..", "include", "wfot_dev.h"),
            os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every translation unit in parallel (nvcc -c), then link libwfot.so."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + (["-Xptxas", "-v"] if verbose else [])
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = [os.path.join(objdir, os.path.basename(s)[:-3] + ".o") for s in srcs]

    def cc(args):
        src, obj = args
        return subprocess.run([_nvcc()] + flags + ["-c", "-o", obj, src], capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        results = list(ex.map(cc, zip(srcs, objs)))
    for r in results:
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
    if any(r.returncode != 0 for r in results):
        raise RuntimeError("nvcc failed building libwfot.so")
    r = subprocess.run([_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed linking libwfot.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
