"""Builds librfid_b200.so (CUDA kernels + C-ABI) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librfid_b200.so")
SOURCES = ["rfid_b200.cu"]
import glob
HEADERS = sorted(os.path.basename(h) for h in glob.glob(os.path.join(CSRC, "*.cuh"))) + ["../../include/rfid_b200.h"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              # no FMA contraction: the reference's x86-64 objects contain none
              "-fmad=false",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-diag-suppress", "550"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile if needed; returns the library path.  Raises if nvcc is unavailable and no library exists."""
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("nvcc not found and %s is missing" % LIB)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
