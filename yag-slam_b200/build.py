"""Builds csrc/libysm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("YSM_LIB") or os.path.join(CSRC, "libysm_b200.so")  # YSM_LIB: A/B another build of the library
SOURCES = ["ysm.cu", "ysm_kernels.cuh", "ysm_resident.cuh", "ysm_internal.h", "ysm_occ.cu", "ysm_chains.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    # cell indices come from Round(double): no FMA contraction, IEEE div/sqrt (defaults)
    "-fmad=false", "-Xcompiler", "-fPIC,-O2", "-shared",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(_HERE, "..", "include", "ysm.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH, "ysm.cu", "ysm_occ.cu", "ysm_chains.cu"]
    subprocess.check_call(cmd, cwd=CSRC)
    return SO_PATH


PYFAST_SO = os.path.join(_HERE, "_ysm_pyfast.so")  # native binding of Wrapper.match_scan (csrc/ysm_pyfast.c)


def build_pyfast(force=False):
    """CPython extension over the C ABI (host glue only: it calls ysm_match_batch, it computes nothing)."""
    import sysconfig
    src = os.path.join(CSRC, "ysm_pyfast.c")
    hdr = os.path.join(_HERE, "..", "include", "ysm.h")
    if not force and os.path.exists(PYFAST_SO) and os.path.getmtime(PYFAST_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return PYFAST_SO
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-fPIC", "-shared", "-Wall", "-I", sysconfig.get_paths()["include"],
           "-o", PYFAST_SO, src]
    subprocess.check_call(cmd, cwd=CSRC)
    return PYFAST_SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_pyfast(force=True))
