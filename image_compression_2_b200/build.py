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
DEPS = SOURCES + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")] + [
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


def build_library(force=False, verbose=False, out=None, defines=()):
    """Compile the CUDA library if it is missing or stale. Returns its path.
    `out`/`defines` build a debug variant elsewhere (tools/dec_profile.py)."""
    target = out or LIB
    if out is None and not force and not needs_build():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build liblatentcodec.so")
    # Several processes may get here at once (every rank under torchrun finds the library stale): one builds, the others
    # wait on the lock and find it fresh; every build writes its own temporary file and renames it into place.
    import fcntl
    with open(target + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if out is None and not force and not needs_build():
                return LIB
            tmp = "%s.%d.tmp" % (target, os.getpid())
            cmd = ([nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) +
                   ["-o", tmp] + SOURCES)
            try:
                subprocess.check_call(cmd)
                os.replace(tmp, target)
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return target


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
