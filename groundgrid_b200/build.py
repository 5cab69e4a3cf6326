"""Builds the native libraries IN-TREE (the built .so travels to the GPU box with the snapshot).

    python -m groundgrid_b200.build          # libgroundgrid_b200.so (+ host mirror)

nvcc cross-compiles sm_100a without a GPU.  --fmad=false: the reference is built for baseline
x86-64 (no FMA), and bit-exact labels need the same roundings.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgroundgrid_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function",
    "-shared",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_core(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("gg_kernels.cu", "gg_capi.cu", "gg_host.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("gg_internal.h", "gg_host.h")] + [os.path.join(ROOT, "include", "groundgrid_b200.h"), __file__]
    if not force and not _stale(LIB, deps):
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


def build_all(force=False, verbose=False):
    out = [build_core(force, verbose)]
    try:
        from . import build_host
        out.append(build_host.build(force))
    except ImportError:
        pass
    return out


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
