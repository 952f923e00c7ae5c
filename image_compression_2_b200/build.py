"""Builds liblatentcodec.so (hand-written CUDA for sm_100a) in-tree with nvcc.

The built library sits next to this file so it travels to the GPU box with the repo snapshot.
`--fmad=false`: the parity path is a sequence of separately rounded IEEE operations
(SURVEY.md section 7), never contracted into FMAs.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblatentcodec.so")
SOURCES = [os.path.join(CSRC, "latentcodec.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "lc_coder.cuh"), os.path.join(CSRC, "lc_common.cuh"),
                  os.path.join(os.path.dirname(HERE), "include", "latentcodec.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_library(force=False, verbose=False):
    """Compile the CUDA library if it is missing or stale. Returns its path."""
    if not force and not needs_build():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build liblatentcodec.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB + ".tmp"] + SOURCES
    subprocess.check_call(cmd)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
