"""Generates the fixtures of this directory:  python tests/golden/make_golden.py

The reference cannot be built or imported in this image, so these vectors do NOT come from
it: they come from tests/pyref.py, the pure-Python restatement written from SURVEY appendix A independently of the
C++ oracle and of the CUDA kernels.  They pin all three implementations against one committed set of numbers
(tests/test_oracle_known_answers.py::test_golden_fixtures for the oracle, tests/test_gpu_parity.py::
test_golden_fixtures_on_gpu for the CUDA path); parity with the real reference stays "unpinned".

Each case_<name>.npz holds, per scan k: points_k (PointXYZIR records as raw bytes), origin_k, base_z_k and the expected
labels_k, order_k; the prior before the first scan (initial ground / groundpatch) and the layers after the last one.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import pyref  # noqa: E402
from groundgrid_b200 import synth  # noqa: E402
from oracle import Oracle  # noqa: E402  (only for the expectedPoints table E, which test_oracle_known_answers pins separately)

CASES = {
    # name: (dimension, resolution, config overrides, scans: (seed, points pushed below ground))
    "n100_default": (33.0, 0.33, {}, [(100, 0), (101, 150), (102, 60)]),
    "n100_small_patch_radius": (33.0, 0.33, {"patch_size_change_distance": 6.0, "point_count_cell_variance_threshold": 4},
                                [(110, 0), (111, 120)]),
    "n61_coarse": (30.5, 0.5, {"ground_patch_detection_minimum_point_count_threshold": 0.4, "outlier_tolerance": 0.05},
                   [(120, 0), (121, 100)]),
}


def small_scan(seed, below_ground):
    scene = synth.make_scene(seed=seed // 10 * 10, n_boxes=10, rmin=4.0, rmax=15.0)
    pts, org = synth.lidar_scan(scene, beams=24, elev_deg=(2.0, -24.8), az_steps=192, seed=seed)
    if below_ground:
        rng = np.random.default_rng(seed + 77)
        idx = rng.choice(len(pts), below_ground, replace=False)
        pts["z"][idx] -= rng.uniform(0.3, 1.0, below_ground).astype(np.float32)
    return pts, org


def main():
    for name, (dim, res, cfg, scans) in CASES.items():
        o = Oracle(dim, res)
        o.init_map(0.0, 0.0, 0.0)
        geo = pyref.Geo(dim, res, 0.0, 0.0)
        assert geo.n == o.n
        E = o.expected_points()
        G = o.layer("ground").copy()      # initGroundGrid: ground = odom z, groundpatch = 1e-7 (GroundGrid.cpp:50-80)
        C = o.layer("groundpatch").copy()
        out = {"dimension": np.float64(dim), "resolution": np.float32(res), "n_scans": np.int32(len(scans)),
               "config_keys": np.array(sorted(cfg), dtype="U64"), "config_values": np.array([cfg[k] for k in sorted(cfg)], np.float64),
               "ground_0": G.copy(), "groundpatch_0": C.copy()}
        for k, (seed, below) in enumerate(scans):
            pts, org = small_scan(seed, below)
            r = pyref.filter_cloud(pts, org, 0.03 * k, G, C, E, geo, cfg)      # advances G, C in place
            out[f"points_{k}"] = np.ascontiguousarray(pts).view(np.uint8).copy()
            out[f"origin_{k}"] = np.asarray(org, np.float32)
            out[f"base_z_{k}"] = np.float64(0.03 * k)
            out[f"labels_{k}"] = np.asarray(r["labels"], np.uint8)
            out[f"order_{k}"] = np.asarray(r["order"], np.uint32)
            print(name, k, len(pts), np.bincount(out[f"labels_{k}"], minlength=100)[[0, 49, 99]], len(r["outliers"]))
        out["ground"] = G.copy()
        out["groundpatch"] = C.copy()
        out["variance"] = np.asarray(r["variance"], np.float32)
        out["points_layer"] = np.asarray(r["points"], np.float32)
        out["minGroundHeight"] = np.asarray(r["minGroundHeight"], np.float32)
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **out)


if __name__ == "__main__":
    main()
