"""Build libkexp.so in-tree with nvcc for sm_100a (B200).  No JIT cache, no fallback."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkexp.so")
SOURCES = ["kexp_api.cu", "kexp_krylov.cu"]
HEADERS = ["kexp_kernels.cuh", "kexp_internal.h", "kexp_krylov_kernels.cuh", os.path.join("..", "..", "include", "kexp.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("KEXP_NVCC_EXTRA", "").split()
    out = os.environ.get("KEXP_LIB_OUT", LIB)
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=CSRC)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
