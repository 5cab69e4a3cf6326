"""The C++ host mirror (groundgrid::GroundGridNodelet -> GroundGrid / GroundSegmentation over the
C-ABI) against the oracle: the nodelet callbacks are driven by a scripted sequence of odometry,
TF and PointCloud2 messages; the published clouds and the terrain layers must be bit-identical."""
import math
import os
import struct
import subprocess

import numpy as np
import pytest

from groundgrid_b200 import build_host, synth
from oracle import Oracle, POINT_DTYPE


def quat_yaw(yaw):
    return (0.0, 0.0, math.sin(yaw / 2.0), math.cos(yaw / 2.0))


def to_matrix(t, q):
    """tf2::Matrix3x3::setRotation order (shim/tf2_ros/transform_listener.h:toMatrix)."""
    f = np.float64
    qx, qy, qz, qw = (f(v) for v in q)
    d = qx * qx + qy * qy + qz * qz + qw * qw
    s = f(2.0) / d
    xs, ys, zs = qx * s, qy * s, qz * s
    wx, wy, wz = qw * xs, qw * ys, qw * zs
    xx, xy, xz = qx * xs, qx * ys, qx * zs
    yy, yz, zz = qy * ys, qy * zs, qz * zs
    return np.array([[1.0 - (yy + zz), xy - wz, xz + wy, t[0]],
                     [xy + wz, 1.0 - (xx + zz), yz - wx, t[1]],
                     [xz - wy, yz + wx, 1.0 - (xx + yy), t[2]]], np.float64)


def test_host_libraries_build_with_the_reference_split_and_export_the_class_symbols():
    """The reference builds three libraries (CMakeLists.txt:96-144): GroundGrid, GroundSegmentation, the nodelet."""
    build_host.build()
    want = {
        build_host.LIB_GRID: ("groundgrid::GroundGrid::update(", "groundgrid::GroundGrid::initGroundGrid(", "groundgrid::GroundGrid::setConfig("),
        build_host.LIB_SEG: ("groundgrid::GroundSegmentation::filter_cloud(", "groundgrid::GroundSegmentation::init(",
                             "groundgrid::GroundSegmentation::setConfig(", "groundgrid::GroundSegmentation::insert_cloud(",
                             "groundgrid::GroundSegmentation::detect_ground_patches(", "groundgrid::GroundSegmentation::spiral_ground_interpolation(",
                             "groundgrid::GroundSegmentation::interpolate_cell(", "void groundgrid::GroundSegmentation::detect_ground_patch<3>(",
                             "void groundgrid::GroundSegmentation::detect_ground_patch<5>("),
        build_host.LIB_NODELET: ("groundgrid::GroundGridNodelet::points_callback(", "groundgrid::GroundGridNodelet::odom_callback(",
                                 "groundgrid::GroundGridNodelet::onInit(", "groundgrid::GroundGridNodelet::callbackReconfigure("),
    }
    for lib, syms in want.items():
        assert os.path.basename(lib) in ("libgroundgrid_lib.so", "libgroundgrid_groundsegmentation_lib.so", "libgroundgrid_nodelet.so")
        out = subprocess.run(["nm", "-DC", "--defined-only", lib], capture_output=True, text=True).stdout
        for sym in syms:
            assert sym in out, (lib, sym)


@pytest.mark.gpu
def test_nodelet_callbacks_match_oracle(tmp_path):
    build_host.build()
    dim, res = 99.0, 0.33
    scene = synth.make_scene(seed=11, stream_len=10.0)
    o = Oracle(dim, res)
    script = [struct.pack("<iff", 4, dim, res)]
    expect = []
    for k in range(4):
        ex, ey, yaw = 1.1 * k, 0.3 * k, 0.02 * k
        frame_is_map = k % 2 == 0
        step = 32 if k < 3 else 18                       # last scan: the KITTI player's 18-byte points
        pts_map, org = synth.scan_64(scene, ego_xy=(ex, ey), yaw=yaw, seed=50 + k)
        pts_base, _ = synth.scan_64(scene, ego_xy=(ex, ey), yaw=yaw, seed=50 + k, frame="base")
        t_mb, q_mb = (ex, ey, 0.0), quat_yaw(yaw)           # map <- base_link
        t_mv, q_mv = (ex, ey, synth.SENSOR_HEIGHT), quat_yaw(yaw)   # map <- velodyne (sensor 1.73 m above base)
        Rb = to_matrix((0, 0, 0), quat_yaw(-yaw))[:, :3]
        t_bm = tuple(-(Rb @ np.array(t_mb)))                # base_link <- map
        q_bm = quat_yaw(-yaw)
        if frame_is_map:
            cloud_in = pts_map
            cloud_map = pts_map
        else:
            # sensor-frame cloud: base-frame points lowered by the sensor height
            cloud_in = pts_base.copy()
            cloud_in["z"] = (pts_base["z"].astype(np.float64) - synth.SENSOR_HEIGHT).astype(np.float32)
            M = to_matrix(t_mv, q_mv)
            x, y, z = (cloud_in[c].astype(np.float64) for c in "xyz")
            cloud_map = cloud_in.copy()
            cloud_map["x"] = ((M[0, 0] * x + M[0, 1] * y + M[0, 2] * z) + M[0, 3]).astype(np.float32)
            cloud_map["y"] = ((M[1, 0] * x + M[1, 1] * y + M[1, 2] * z) + M[1, 3]).astype(np.float32)
            cloud_map["z"] = ((M[2, 0] * x + M[2, 1] * y + M[2, 2] * z) + M[2, 3]).astype(np.float32)
        if step == 18:
            raw = np.zeros(len(cloud_in), np.dtype({"names": ["x", "y", "z", "intensity", "ring"], "formats": ["<f4"] * 4 + ["<u2"],
                                                   "offsets": [0, 4, 8, 12, 16], "itemsize": 18}))
            for c in ("x", "y", "z", "intensity", "ring"):
                raw[c] = cloud_in[c]
            payload = raw.tobytes()
        else:
            payload = np.ascontiguousarray(cloud_in).tobytes()
        script.append(struct.pack("<3d7d7d7d3i", ex, ey, 0.0, *t_mb, *q_mb, *t_mv, *q_mv, *t_bm, *q_bm, int(frame_is_map), step, len(cloud_in)))
        script.append(payload)
        # oracle: update (except first), origin = T(map<-velodyne) * 0, base_z = translation z of map<-base
        T_bm = to_matrix(t_bm, q_bm)
        if k == 0:
            o.init_map(ex, ey, 0.0)
        else:
            o.update(ex, ey, T_bm)
        origin = np.array([np.float32(t_mv[0]), np.float32(t_mv[1]), np.float32(t_mv[2])], np.float32)
        # fromROSMsg zero-fills padding and copies x,y,z,intensity,ring
        clean = np.zeros(len(cloud_map), POINT_DTYPE)
        for c in ("x", "y", "z", "intensity", "ring"):
            clean[c] = cloud_map[c]
        _, _, out_cloud = o.filter_cloud(clean, origin, t_mb[2], threads=1, want_cloud=True)
        expect.append((out_cloud.tobytes(), o.layer("ground").copy(), o.layer("groundpatch").copy()))
    sp, op = tmp_path / "script.bin", tmp_path / "out.bin"
    sp.write_bytes(b"".join(script))
    r = subprocess.run([build_host.TEST_BIN, str(sp), str(op)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host mirror ok" in r.stdout
    data = op.read_bytes()
    pos = 0
    n = o.n
    for k, (cloud_bytes, G, C) in enumerate(expect):
        (n_out,) = struct.unpack_from("<i", data, pos)
        pos += 4
        got = data[pos:pos + n_out * 32]
        pos += n_out * 32
        assert n_out * 32 == len(cloud_bytes), f"scan {k}: {n_out} points published, oracle {len(cloud_bytes) // 32}"
        assert got == cloud_bytes, f"scan {k}: published cloud differs"
        Gg = np.frombuffer(data, np.float32, n * n, pos).reshape(n, n, order="F")
        pos += 4 * n * n
        Cg = np.frombuffer(data, np.float32, n * n, pos).reshape(n, n, order="F")
        pos += 4 * n * n
        assert np.array_equal(Gg, G) and np.array_equal(Cg, C), f"scan {k}: terrain layers differ"
