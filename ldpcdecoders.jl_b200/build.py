"""Build libldpcb200.so in-tree with nvcc for sm_100a (called by __graft_entry__.build())."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libldpcb200.so")
SOURCES = ["ldpcb200.cu"]
HEADERS = ["bp_math.cuh", "bp_kernel.cuh", "formats.cuh", "../../include/ldpcb200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # never contract a*b+c: every FP64 op of the parity path is rounded separately
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(SRC_DIR, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + [os.path.join(SRC_DIR, f) for f in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (see %s)" % os.path.join(LIB_DIR, "build.log"))
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
