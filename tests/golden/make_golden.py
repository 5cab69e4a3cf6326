"""Generates the fixtures of this directory FROM THE REFERENCE ITSELF:  python tests/golden/make_golden.py

The vectors come from oracle/_ref/libgg_ref.so, i.e. the unmodified /root/reference/src/GroundSegmentation.cpp and
GroundGrid.cpp compiled against the CPU stand-ins of oracle/ref_shim/ (oracle/build_ref.py), run at thread_count = 1
(the shipped thread_count = 8 races on the shared matrices and is not reproducible).  /root/reference only exists in
the build container, so the vectors are committed; the oracle port (-m "not gpu") and the CUDA path (-m gpu) are both
checked against them bit for bit (tests/test_golden.py).

Each case_<name>.npz holds: geometry and config; the map right after creation (ground_0 / groundpatch_0); per scan k
the odometry pose and the base_link<-map transform handed to GroundGrid::update (quaternion + translation as the ROS
message carries it, and the row-major 3x4 matrix tf2 derives from it), whether the map moved, the map position, the
rolled prior (ground / groundpatch before the scan), the cloud (PointXYZIR records as raw bytes), cloudOrigin, the z
of mapToBase, and the reference's answer: label per input point (0 = absent from the output cloud), output order;
after the last scan all eleven layers.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from groundgrid_b200 import synth  # noqa: E402
from oracle import LAYERS  # noqa: E402
from oracle.ref import Reference, base_from_map_qt, tf2_matrix  # noqa: E402

CASES = {
    # name: (dimension, resolution, config overrides, scans: (seed, points pushed below ground, ego step in m, pitch))
    "n100_default": (33.0, 0.33, {}, [(100, 0, 0.0, 0.0), (101, 150, 0.0, 0.0), (102, 60, 0.0, 0.0)]),
    "n100_small_patch_radius": (33.0, 0.33, {"patch_size_change_distance": 6.0, "point_count_cell_variance_threshold": 4},
                                [(110, 0, 0.0, 0.0), (111, 120, 0.0, 0.0)]),
    "n55_coarse_odd": (33.0, 0.6, {"ground_patch_detection_minimum_point_count_threshold": 0.4, "outlier_tolerance": 0.05},
                       [(120, 0, 0.0, 0.0), (121, 100, 0.0, 0.0)]),
    "n100_rolling": (33.0, 0.33, {}, [(130, 0, 0.0, 0.0), (131, 80, 0.9, 0.02), (132, 80, 0.2, 0.02), (133, 80, 1.4, -0.03)]),
    "n61_odd_rolling": (20.0, 0.33, {"occupied_cells_decrease_factor": 3.0, "max_ring": 20},
                        [(140, 0, 0.0, 0.0), (141, 60, -0.7, 0.01), (142, 60, -0.7, 0.01)]),
}


def small_scan(seed, below_ground, ego_xy, yaw):
    scene = synth.make_scene(seed=seed // 10 * 10, n_boxes=10, rmin=4.0, rmax=15.0)
    pts, org = synth.lidar_scan(scene, ego_xy=ego_xy, yaw=yaw, beams=24, elev_deg=(2.0, -24.8), az_steps=192, seed=seed)
    if below_ground:
        rng = np.random.default_rng(seed + 77)
        idx = rng.choice(len(pts), below_ground, replace=False)
        pts["z"][idx] -= rng.uniform(0.3, 1.0, below_ground).astype(np.float32)
    return pts, org


def main():
    for name, (dim, res, cfg, scans) in CASES.items():
        r = Reference(dim, res)
        if cfg:
            r.set_config(**cfg)
        r.init_map(0.0, 0.0, 0.0)
        out = {"dimension": np.float64(dim), "resolution": np.float32(res), "cells": np.int32(r.n), "n_scans": np.int32(len(scans)),
               "config_keys": np.array(sorted(cfg), dtype="U64"), "config_values": np.array([cfg[k] for k in sorted(cfg)], np.float64),
               "ground_0": r.layer("ground"), "groundpatch_0": r.layer("groundpatch"), "expected": r.expected_points()}
        ex = ey = 0.0
        for k, (seed, below, step, pitch) in enumerate(scans):
            ex += step
            ey -= 0.4 * step
            yaw = 0.05 * k
            q, t = base_from_map_qt(ex, ey, yaw, 0.03 * k, pitch)
            moved = r.update(ex, ey, q, t) if k else 0
            pts, org = small_scan(seed, below, (ex, ey), yaw)
            out[f"pose_{k}"] = np.array([ex, ey], np.float64)
            out[f"q_{k}"], out[f"t_{k}"], out[f"T_{k}"] = q, t, tf2_matrix(q, t)
            out[f"moved_{k}"] = np.int32(moved)
            out[f"position_{k}"] = r.position()
            out[f"prior_ground_{k}"], out[f"prior_groundpatch_{k}"] = r.layer("ground"), r.layer("groundpatch")
            labels, order, _ = r.filter_cloud(pts, org, 0.03 * k)
            out[f"points_{k}"] = np.ascontiguousarray(pts).view(np.uint8).copy()
            out[f"origin_{k}"] = np.asarray(org, np.float32)
            out[f"base_z_{k}"] = np.float64(0.03 * k)
            out[f"labels_{k}"] = labels
            out[f"order_{k}"] = order
            n_out = int(r.layer("pointsRaw").sum()) - int((labels != 0).sum())
            print(name, k, len(pts), "moved", moved, "absent/49/99", np.bincount(labels, minlength=100)[[0, 49, 99]], "inside-but-absent", n_out)
        for lname in LAYERS:
            out["layer_" + lname] = r.layer(lname)
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **out)


if __name__ == "__main__":
    for f in os.listdir(HERE):
        if f.startswith("case_") and f.endswith(".npz"):
            os.remove(os.path.join(HERE, f))
    main()
