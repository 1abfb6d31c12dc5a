"""Build libkexp.so in-tree with nvcc for sm_100a (B200).  No JIT cache, no fallback.
Every translation unit is compiled in its own nvcc process (in parallel), then linked."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkexp.so")
OBJDIR = os.path.join(CSRC, "build")
SOURCES = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "kexp.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


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
    tag = os.path.splitext(os.path.basename(out))[0]
    os.makedirs(OBJDIR, exist_ok=True)
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJDIR, f"{tag}_{os.path.splitext(src)[0]}.o")
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            cmd += ["-Xptxas", "-v"]
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, cwd=CSRC)))
        objs.append(obj)
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ["-ldl"], check=True, cwd=CSRC)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
