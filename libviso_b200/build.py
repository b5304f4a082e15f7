"""Build libviso_b200.so (the C-ABI library: hand-written sm_100a kernels + host glue) in-tree with nvcc.

  python -m libviso_b200.build [--force]

-fmad=false: the FP64 estimation kernels must evaluate the reference's expressions with separate multiplies and
adds (what a stock x86-64 build of the reference does) so Jacobians, pivots and inlier decisions agree bit for bit.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libviso_b200.so")
SOURCES = ["detect.cu", "match.cu", "sort_circle.cu", "estimation.cu", "geometry.cu", "capi.cu", "capi_seq.cu"]
HEADERS = [os.path.join(CSRC, "viso_dev.h"), os.path.join(CSRC, "introsort.h"), os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "capi_internal.h"), os.path.join(CSRC, "glibc_sincos.h"), os.path.join(CSRC, "glibc_sincostab.inc"),
           os.path.join(os.path.dirname(HERE), "include", "viso_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fno-builtin-sin,-fno-builtin-cos", "-shared"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: cannot build libviso_b200.so")
    return p


def up_to_date():
    if not os.path.exists(SO):
        return False
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return SO
    cmd = [nvcc_path()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", SO]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
