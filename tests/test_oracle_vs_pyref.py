"""The C++ oracle against the independent pure-Python restatement (tests/pyref.py): every layer
and every label must agree bit for bit on small seeded scans, including the outlier ray-march
(needs a confident prior), the 5x5 patch branch and a rolled prior."""
import numpy as np
import pytest

import pyref
from groundgrid_b200 import synth
from oracle import Oracle


def small_scan(seed, below_ground=0, ego=(0.0, 0.0)):
    scene = synth.make_scene(seed=seed, n_boxes=10, rmin=4.0, rmax=15.0)
    pts, org = synth.lidar_scan(scene, ego_xy=ego, beams=24, elev_deg=(2.0, -24.8), az_steps=192, seed=seed)
    if below_ground:
        rng = np.random.default_rng(seed + 77)
        idx = rng.choice(len(pts), below_ground, replace=False)
        pts["z"][idx] -= rng.uniform(0.3, 1.0, below_ground).astype(np.float32)
    return pts, org


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("patch_dist", [20.0, 6.0])
def test_two_scan_sequence_bit_exact(patch_dist):
    dim, res = 33.0, 0.33
    o = Oracle(dim, res)
    o.set_config(patch_size_change_distance=patch_dist)
    o.init_map(0.0, 0.0, 0.0)
    N = o.n
    geo = pyref.Geo(dim, res, 0.0, 0.0)
    assert geo.n == N == 100
    E = o.expected_points()
    G = o.layer("ground").copy()
    C = o.layer("groundpatch").copy()
    cfg = dict(patch_size_change_distance=patch_dist)

    n_outliers = 0
    for scan in range(2):
        pts, org = small_scan(100 + scan, below_ground=(0 if scan == 0 else 150))
        for stage in (1, 2, 3):
            # run the oracle up to `stage` on a copy of the prior and compare with pyref
            o.set_layer("ground", G)
            o.set_layer("groundpatch", C)
            o.filter_cloud(pts, org, 0.05, threads=1, stop_after=stage)
            Gp, Cp = G.copy(), C.copy()
            r = pyref.filter_cloud(pts, org, 0.05, Gp, Cp, E, geo, cfg, stop_after=stage)
            assert same(o.layer("points"), r["count"]), (scan, stage)
            for name in ("m2", "minGroundHeight", "maxGroundHeight", "meanVariance", "groundCandidates", "planeDist", "pointsRaw"):
                assert same(o.layer(name), r[name]), (scan, stage, name)
            if stage >= 2:
                assert same(o.layer("variance"), r["variance"]), (scan, stage)
            assert same(o.layer("ground"), Gp), (scan, stage)
            assert same(o.layer("groundpatch"), Cp), (scan, stage)
        o.set_layer("ground", G)
        o.set_layer("groundpatch", C)
        labels, order, _ = o.filter_cloud(pts, org, 0.05, threads=1)
        r = pyref.filter_cloud(pts, org, 0.05, G, C, E, geo, cfg)      # advances G, C in place
        assert same(labels, r["labels"])
        assert same(order, r["order"])
        assert same(o.layer("points"), r["points"])
        assert same(o.layer("ground"), G) and same(o.layer("groundpatch"), C)
        n_outliers += len(r["outliers"])
        assert (labels == 99).sum() > 50 and (labels == 49).sum() > 500
    assert n_outliers > 0, "second scan must exercise the outlier ray-march"
