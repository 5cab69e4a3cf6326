"""The steps next to the hot path: device unpack + transform of PointCloud2
payloads (f1), terrain image export (f3), on-device evaluation tallies (f4) -- against the numpy
restatements in oracle/nextrows.py.  f2 (dead layers) is covered by tests/test_gpu_parity.py."""
import numpy as np
import pytest

from groundgrid_b200 import capi, evalmetrics, synth
from oracle import Oracle, nextrows

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("step,offsets", [(32, (0, 4, 8, 16, 20)), (18, (0, 4, 8, 12, 16)), (22, (4, 8, 12, -1, 20))])
@pytest.mark.parametrize("in_map_frame", [False, True])
def test_f1_unpack_transform_then_filter(step, offsets, in_map_frame):
    dim, res = 99.0, 0.33
    scene = synth.make_scene(seed=21)
    ego, yaw = (3.0, -1.5), 0.3
    pts_map, org = synth.scan_64(scene, ego_xy=ego, yaw=yaw, seed=21)
    pts_base, _ = synth.scan_64(scene, ego_xy=ego, yaw=yaw, seed=21, frame="base")
    src = pts_map if in_map_frame else pts_base
    n = len(src)
    raw = np.zeros((n, step), np.uint8)
    for name, off, width in (("x", offsets[0], 4), ("y", offsets[1], 4), ("z", offsets[2], 4), ("intensity", offsets[3], 4), ("ring", offsets[4], 2)):
        if off >= 0:
            raw[:, off:off + width] = np.ascontiguousarray(src[name]).view(np.uint8).reshape(n, width)
    T = None
    if not in_map_frame:
        c, s = np.cos(yaw), np.sin(yaw)
        T = np.array([[c, -s, 0.0, ego[0]], [s, c, 0.0, ego[1]], [0.0, 0.0, 1.0, 0.0]], np.float64)   # map <- base
    want_cloud = nextrows.unpack_transform(raw, n, step, offsets, T)
    g = capi.GroundGridB200(dim, res, n_slots=1, max_points=140000, full_layers=True)
    o = Oracle(dim, res)
    g.init_map(ego[0], ego[1], 0.0)
    o.init_map(ego[0], ego[1], 0.0)
    keep = g.upload_cloud_msg(raw, n, step, offsets, T)
    d = g.make_descs([0], [n], [org], [0.0])
    g.run_scans(d)
    labels = g.download_labels(n)
    g.synchronize()
    index, cloud = g.get_output(want_cloud=True)
    lab_o, idx_o, cloud_o = o.filter_cloud(want_cloud, org, 0.0, threads=1, want_cloud=True)
    assert np.array_equal(labels, lab_o)
    assert np.array_equal(index, idx_o)
    assert cloud.tobytes() == cloud_o.tobytes()          # x, y, z (transformed), ring and intensity survive the unpack
    del keep


def test_f3_terrain_image_and_f4_eval_counts():
    dim, res = 120.0, 0.33
    scene = synth.make_scene(seed=31)
    g = capi.GroundGridB200(dim, res, n_slots=1, max_points=140000, full_layers=True)
    o = Oracle(dim, res)
    g.init_map(0.0, 0.0, 0.0)
    o.init_map(0.0, 0.0, 0.0)
    rng = np.random.default_rng(1)
    ids = np.array(sorted(evalmetrics.LABELS), np.uint16)
    total = np.zeros((1024, 2), np.uint64)
    for k in range(3):
        pts, org = synth.scan_64(scene, ego_xy=(0.5 * k, 0.0), seed=31 + k)
        # ground truth label in `ring` (scripts/kitti_data_publisher.py:122-132): flat points are road/terrain, the rest objects
        gt = np.where(pts["z"] < 0.15, rng.choice([40, 48, 72, 49, 60], len(pts)), rng.choice([10, 50, 71, 80, 30, 70, 0], len(pts)))
        pts["ring"] = gt.astype(np.uint16)
        if k:
            T = synth.base_from_map(0.5 * k, 0.0)
            g.update_pose(0.5 * k, 0.0, T)
            o.update(0.5 * k, 0.0, T)
        labels = g.filter_cloud(pts, org, 0.0)
        lab_o, _, _ = o.filter_cloud(pts, org, 0.0, threads=1)
        assert np.array_equal(labels, lab_o)
        g.eval_accumulate()
        total += nextrows.eval_counts(lab_o, pts["ring"])
    counts = g.eval_read(reset=True)
    assert np.array_equal(counts, total)
    assert np.all(g.eval_read() == 0)
    m = evalmetrics.metrics(counts)
    assert 0.9 < m["precision"] <= 1.0 and 0.9 < m["recall"] <= 1.0 and 0.8 < m["iou_ground"] <= 1.0
    img = g.terrain_image()
    want = nextrows.terrain_image(o.layer("ground"), o.layer("pointsRaw"))
    assert np.array_equal(img, want)
    assert img[:, :, 1].sum() > 100
    assert ids.max() < 1024


def test_layer_images_8bit():
    """f3: the 8UC1 image toImage<unsigned char, 1> hands to cv::applyColorMap, for live and dead layers, and with the
    NaN strips a map roll leaves in per-scan layers."""
    g = capi.GroundGridB200(33.0, 0.33, n_slots=2, max_points=20000, full_layers=True)
    o = Oracle(33.0, 0.33)
    scene = synth.make_scene(seed=5, n_boxes=8, rmin=4.0, rmax=14.0)
    pts, org = synth.lidar_scan(scene, beams=32, az_steps=512, seed=5)
    for slot in (0, 1):
        g.init_map(0.0, 0.0, 0.3 * slot, slot=slot)
    o.init_map(0.0, 0.0, 0.3)
    g.filter_cloud(pts, org, 0.0, slot=1)
    o.filter_cloud(pts, org, 0.0, threads=1)
    for name in ("ground", "groundpatch", "variance", "points", "pointsRaw", "meanVariance"):
        img, lo, hi = g.layer_image_u8(name, slot=1)
        want, wlo, whi = nextrows.layer_image_u8(o.layer(name))
        assert (lo, hi) == (wlo, whi), name
        assert np.array_equal(img, want), (name, int((img != want).sum()))
    g.close()


def test_per_phase_entry_points_match_the_pipeline():
    """The reference's public per-phase methods (GroundSegmentation.h:56-62) through the C-ABI: insert_cloud (stage 1 +
    index lists), detect_ground_patches, spiral_ground_interpolation run one by one give the layers of the fused
    pipeline, and the single-cell calls match the oracle's."""
    g = capi.GroundGridB200(33.0, 0.33, n_slots=1, max_points=20000, full_layers=True)
    o = Oracle(33.0, 0.33)
    scene = synth.make_scene(seed=9, n_boxes=8, rmin=4.0, rmax=14.0)
    g.init_map(0.0, 0.0, 0.0)
    o.init_map(0.0, 0.0, 0.0)
    for k in range(2):
        pts, org = synth.lidar_scan(scene, beams=32, az_steps=512, seed=90 + k)
        if k:
            pts["z"][::37] -= 0.7
        # phase by phase
        g.run_single(pts, org, 0.25, stop_after=1)
        codes = g.point_classes(len(pts))
        g.detect_ground_patches()
        g.spiral_ground_interpolation(0.25)
        g.synchronize()
        lab_o, _, _ = o.filter_cloud(pts, org, 0.25, threads=1, stop_after=3)
        for name in ("ground", "groundpatch", "variance", "minGroundHeight"):
            assert np.array_equal(g.layer(name), o.layer(name)), (k, name)
        # the index lists of insert_cloud: classes against the labels of a full oracle run on a twin map
        cls = codes >> 24
        assert set(np.unique(cls)) <= {0, 1, 2, 3, 4, 5}
        assert (cls == 5).sum() > 0 if k else True
    # single cells
    rng = np.random.default_rng(3)
    G = rng.normal(0.0, 0.4, (g.n, g.n)).astype(np.float32)
    Cf = rng.uniform(0.0, 1.0, (g.n, g.n)).astype(np.float32)
    for m in (g, o):
        m.set_layer("ground", G)
        m.set_layer("groundpatch", Cf)
    for x, y in ((5, 7), (48, 49), (49, 49), (1, 1), (97, 97)):
        g.interpolate_cell(x, y)
        o.interpolate_cell(x, y)
    g.synchronize()
    assert np.array_equal(g.layer("ground"), o.layer("ground")) and np.array_equal(g.layer("groundpatch"), o.layer("groundpatch"))
    g.close()


def test_upload_cloud_msg_slots_sharing_a_stream_keep_their_own_payload():
    """Eight slots on four streams: two slots share a stream and its payload staging buffer.  All eight PointCloud2
    payloads are uploaded back to back (no sync in between), then scanned; every slot must have unpacked ITS cloud."""
    dim, res = 33.0, 0.33
    g = capi.GroundGridB200(dim, res, n_slots=8, max_points=16384, full_layers=False)
    scene = synth.make_scene(seed=3, n_boxes=8, rmin=4.0, rmax=14.0)
    clouds, raws = [], []
    for s_ in range(8):
        pts, org = synth.lidar_scan(scene, beams=24 + s_, az_steps=256, seed=50 + s_)
        n = len(pts)
        raw = np.zeros((n, 18), np.uint8)
        for name, off, width in (("x", 0, 4), ("y", 4, 4), ("z", 8, 4), ("intensity", 12, 4), ("ring", 16, 2)):
            raw[:, off:off + width] = np.ascontiguousarray(pts[name]).view(np.uint8).reshape(n, width)
        clouds.append((pts, org))
        raws.append(raw)
        g.init_map(0.0, 0.0, 0.0, slot=s_)
    keep = [g.upload_cloud_msg(raws[s_], len(raws[s_]), 18, (0, 4, 8, 12, 16), slot=s_) for s_ in range(8)]
    d = g.make_descs(list(range(8)), [len(c[0]) for c in clouds], [c[1] for c in clouds], [0.0] * 8)
    g.run_scans(d)
    g.synchronize()
    for s_ in range(8):
        pts, org = clouds[s_]
        o = Oracle(dim, res)
        o.init_map(0.0, 0.0, 0.0)
        lab_o, _, _ = o.filter_cloud(pts, org, 0.0, threads=1)
        got = g.download_labels(len(pts), s_)
        g.synchronize()
        assert np.array_equal(got, lab_o), s_
    del keep
    g.close()
