"""Builds hesaff_b200/libhesaff_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m hesaff_b200.build [--force] [-v]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# HESAFF_LIB selects another build of the same library (tuning experiments: tools/build_variants.py)
LIB = os.environ.get("HESAFF_LIB") or os.path.join(HERE, "libhesaff_b200.so")
SOURCES = ["api.cu", "pyramid.cu", "blur_tma.cu", "keypoints.cu", "describe.cu", "describe_large.cu", "match.cu", "export.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "describe.cuh"), os.path.join(HERE, "..", "include", "hesaff_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # the reference is built without FMA contraction; fused multiply-adds appear only where written
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-shared", "-cudart", "static",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
