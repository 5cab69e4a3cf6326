"""Builds oracle/_ref/libgg_ref.so: the UNMODIFIED reference sources compiled against CPU stand-ins.

TEST INFRASTRUCTURE ONLY.  Compiles, from where they lie under /root/reference and with zero edits,

    src/GroundSegmentation.cpp   src/GroundGrid.cpp   (+ include/groundgrid/*.h, include/velodyne_pointcloud/point_types.h)

with g++ directly (the reference's own catkin/CMake build is not run: it needs roscpp, Eigen3, grid_map, PCL,
tf2 and dynamic_reconfigure, none of which is in the image) against oracle/ref_shim/ and the C harness
oracle/ref_harness.cpp.  Flags follow the reference's Release build (CMakeLists.txt:5 "-O3", C++17, baseline
x86-64, no fast-math => no FMA contraction).  Outputs go to oracle/_ref/ only (git-ignored, not
gpurun-ignored: the GPU box has no /root/reference and uses the prebuilt library).  The sha256 of every
reference file that entered the build is written to oracle/_ref/SOURCES.sha256 and compared with the
committed oracle/ref_sources.sha256, so a changed (or edited) reference is noticed.

    python oracle/build_ref.py [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libgg_ref.so")
REF_FILES = [
    "src/GroundSegmentation.cpp",
    "src/GroundGrid.cpp",
    "include/groundgrid/GroundSegmentation.h",
    "include/groundgrid/GroundGrid.h",
    "include/groundgrid/GroundGridFwd.h",
    "include/velodyne_pointcloud/point_types.h",
    "cfg/GroundGrid.cfg",
]
# -fno-gnu-unique: statics of template instantiations (detect_ground_patch<S>, GroundSegmentation.cpp:345-351) must stay
# private to each loaded copy of the library (oracle/ref.py loads one copy per map)
CXXFLAGS = ["-O3", "-DNDEBUG", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-w", "-fno-gnu-unique"]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def _shim_files():
    out = []
    for root, _, files in os.walk(os.path.join(HERE, "ref_shim")):
        out += [os.path.join(root, f) for f in files]
    return out


def build(force=False):
    if not os.path.isdir(REF):
        if os.path.exists(LIB):
            return LIB  # GPU box: prebuilt library travels with the snapshot
        raise RuntimeError("oracle/_ref: /root/reference is absent and no prebuilt library exists")
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(REF, f) for f in REF_FILES]
    deps = srcs + _shim_files() + [os.path.join(HERE, "ref_harness.cpp"), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    sums = "".join(f"{_sha(os.path.join(REF, f))}  {f}\n" for f in REF_FILES)
    pinned = os.path.join(HERE, "ref_sources.sha256")
    if os.path.exists(pinned):
        if open(pinned).read() != sums:
            raise RuntimeError("oracle/_ref: reference sources differ from oracle/ref_sources.sha256")
    else:
        with open(pinned, "w") as f:
            f.write(sums)
    cmd = ["g++"] + CXXFLAGS + ["-I", os.path.join(HERE, "ref_shim"), "-I", os.path.join(REF, "include"), "-o", LIB,
                                os.path.join(REF, "src/GroundSegmentation.cpp"), os.path.join(REF, "src/GroundGrid.cpp"),
                                os.path.join(HERE, "ref_harness.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("oracle/_ref build failed: " + " ".join(cmd))
    with open(os.path.join(OUT, "SOURCES.sha256"), "w") as f:
        f.write(sums)
    with open(os.path.join(OUT, "BUILD.txt"), "w") as f:
        f.write(" ".join(cmd) + "\n" + subprocess.run(["g++", "--version"], capture_output=True, text=True).stdout.splitlines()[0] + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
