"""Builds the C++ host mirror (libgroundgrid_b200_host.so: groundgrid::GroundGrid,
groundgrid::GroundSegmentation, nodelet callbacks) and its test driver.  g++ only."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libgroundgrid_b200_host.so")
TEST_BIN = os.path.join(HOST, "test", "host_mirror_test")
CORE = os.path.join(HERE, "libgroundgrid_b200.so")


def _sources():
    src = [os.path.join(HOST, "src", f) for f in ("GroundGrid.cpp", "GroundSegmentation.cpp", "GroundGridNodelet.cpp")]
    deps = list(src)
    for base, _, files in os.walk(HOST):
        deps += [os.path.join(base, f) for f in files if f.endswith((".h", ".hpp", ".cpp"))]
    return src, deps + [os.path.join(ROOT, "include", "groundgrid_b200.h"), __file__]


def build(force=False):
    src, deps = _sources()
    inc = ["-I", os.path.join(HOST, "include"), "-I", os.path.join(HOST, "shim"), "-I", os.path.join(ROOT, "include")]
    flags = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall", "-Wno-unused-parameter"]
    stale = force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps)
    if stale:
        cmd = ["g++"] + flags + inc + ["-shared", "-o", LIB] + src + ["-L", HERE, "-lgroundgrid_b200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("host mirror build failed")
        cmd = ["g++"] + flags + inc + ["-o", TEST_BIN, os.path.join(HOST, "test", "host_mirror_test.cpp"), "-L", HERE,
                                      "-lgroundgrid_b200_host", "-lgroundgrid_b200", "-Wl,-rpath," + HERE]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("host mirror test build failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
