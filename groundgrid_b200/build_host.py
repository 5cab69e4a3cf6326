"""Builds the C++ host mirror with the reference's library split (CMakeLists.txt:96-144 there) and its test driver:

    libgroundgrid_lib.so                      groundgrid::GroundGrid                 (src/GroundGrid.cpp)
    libgroundgrid_groundsegmentation_lib.so   groundgrid::GroundSegmentation         (src/GroundSegmentation.cpp)
    libgroundgrid_nodelet.so                  groundgrid::GroundGridNodelet          (src/GroundGridNodelet.cpp, links the two above)

all three over libgroundgrid_b200.so (the C-ABI).  g++ only."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HOST = os.path.join(HERE, "host")
LIB_GRID = os.path.join(HERE, "libgroundgrid_lib.so")
LIB_SEG = os.path.join(HERE, "libgroundgrid_groundsegmentation_lib.so")
LIB_NODELET = os.path.join(HERE, "libgroundgrid_nodelet.so")
LIBS = (LIB_GRID, LIB_SEG, LIB_NODELET)
LIB = LIB_NODELET
TEST_BIN = os.path.join(HOST, "test", "host_mirror_test")
CORE = os.path.join(HERE, "libgroundgrid_b200.so")


def _deps():
    deps = []
    for base, _, files in os.walk(HOST):
        deps += [os.path.join(base, f) for f in files if f.endswith((".h", ".hpp", ".cpp"))]
    return deps + [os.path.join(ROOT, "include", "groundgrid_b200.h"), __file__]


def _run(cmd, what):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(what + " failed")


def build(force=False):
    deps = _deps()
    inc = ["-I", os.path.join(HOST, "include"), "-I", os.path.join(HOST, "shim"), "-I", os.path.join(ROOT, "include")]
    flags = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall", "-Wno-unused-parameter"]
    targets = list(LIBS) + [TEST_BIN]
    stale = force or any(not os.path.exists(t) for t in targets) or any(os.path.getmtime(d) > min(os.path.getmtime(t) for t in targets) for d in deps)
    if stale:
        core = ["-L", HERE, "-lgroundgrid_b200", "-Wl,-rpath,$ORIGIN"]
        _run(["g++"] + flags + inc + ["-shared", "-o", LIB_GRID, os.path.join(HOST, "src", "GroundGrid.cpp")] + core, "libgroundgrid_lib")
        _run(["g++"] + flags + inc + ["-shared", "-o", LIB_SEG, os.path.join(HOST, "src", "GroundSegmentation.cpp")] + core, "libgroundgrid_groundsegmentation_lib")
        _run(["g++"] + flags + inc + ["-shared", "-o", LIB_NODELET, os.path.join(HOST, "src", "GroundGridNodelet.cpp"), "-L", HERE, "-lgroundgrid_lib",
                                      "-lgroundgrid_groundsegmentation_lib", "-lgroundgrid_b200", "-Wl,-rpath,$ORIGIN"], "libgroundgrid_nodelet")
        _run(["g++"] + flags + inc + ["-o", TEST_BIN, os.path.join(HOST, "test", "host_mirror_test.cpp"), "-L", HERE, "-lgroundgrid_nodelet",
                                      "-lgroundgrid_lib", "-lgroundgrid_groundsegmentation_lib", "-lgroundgrid_b200", "-Wl,-rpath," + HERE], "host mirror test")
        old = os.path.join(HERE, "libgroundgrid_b200_host.so")
        if os.path.exists(old):
            os.remove(old)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
