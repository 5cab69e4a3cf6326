"""CPU-side checks of the product library: it loads, exports every symbol the header declares,
refuses to compute without a GPU, and its host-side logic (expectedPoints table, map-move
arithmetic, spiral wavefront schedule) agrees with the oracle."""
import os
import re

import numpy as np
import pytest

from groundgrid_b200 import capi
from oracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32
FLT_MIN = f32(np.finfo(np.float32).tiny)


def header_symbols():
    text = open(os.path.join(ROOT, "include", "groundgrid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gg_[a-z_0-9]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = capi.load()
    names = header_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/groundgrid_b200.h but not exported"


def test_struct_layouts_match_the_header():
    import ctypes as C

    assert C.sizeof(capi.Config) == 8 + 11 * 8 + 8        # 2 ints, 11 doubles, int + padding
    assert C.sizeof(capi.ScanDesc) == 40
    assert capi.POINT_DTYPE.itemsize == 32
    cfg = capi.Config()
    capi.load().gg_default_config(C.byref(cfg))
    assert (cfg.point_count_cell_variance_threshold, cfg.max_ring, cfg.thread_count) == (10, 1024, 8)
    assert (cfg.distance_factor, cfg.minimum_distance_factor, cfg.patch_size_change_distance) == (0.0001, 0.0005, 20.0)
    assert (cfg.occupied_cells_decrease_factor, cfg.occupied_cells_point_count_factor) == (5.0, 20.0)
    assert cfg.min_outlier_detection_ground_confidence == 1.25 and cfg.outlier_tolerance == 0.1


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is for machines without one")
    with pytest.raises(capi.GroundGridError) as e:
        capi.GroundGridB200(99.0, 0.33)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("dim,res,n", [(120.0, 0.33, 364), (99.0, 0.33, 300), (120.0, 0.2, 600), (33.0, 0.33, 100)])
def test_expected_points_table_matches_oracle(dim, res, n):
    assert capi.host_cells_per_side(dim, res) == n
    E = capi.host_expected_points(dim, res)
    assert np.array_equal(E, Oracle(dim, res).expected_points())


def test_move_map_matches_oracle():
    rng = np.random.default_rng(9)
    res = float(np.float32(0.33))
    o = Oracle(33.0, 0.33)
    o.init_map(1.5, -2.5, 0.0)
    pos = o.position().copy()
    T = np.eye(4)[:3]
    for _ in range(200):
        target = pos + rng.uniform(-3, 3, 2) * res * rng.choice([0.1, 1.0, 5.0])
        moved, new_pos, shift = capi.host_move_map(res, pos, target)
        A = (np.arange(o.n)[:, None] * o.n + np.arange(o.n)[None, :]).astype(np.float32)
        o.set_layer("ground", A)
        assert o.update(target[0], target[1], T) == int(moved)
        assert np.array_equal(o.position(), new_pos)
        if moved:
            # oracle content shift must equal the reported index shift: new(r,c) = old(r+si, c+sj)
            Gn = o.layer("ground")
            r = np.arange(o.n)[:, None] + shift[0]
            c = np.arange(o.n)[None, :] + shift[1]
            ok = (r >= 0) & (r < o.n) & (c >= 0) & (c < o.n)
            want = (r * o.n + c).astype(np.float32)
            assert np.array_equal(Gn[ok], want[ok])
            assert np.all(o.layer("groundpatch")[~ok] == 0.0)
        pos = new_pos


def _tree9(v):
    return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + (v[7] + v[8])))


def run_level_schedule(G, C, base_z, level_start, visits, res_f, dec_factor=5.0):
    """Executes the spiral as the GPU does: level by level, all visits of a level read first, then write."""
    n = G.shape[0]
    c = n // 2 - 1
    G = G.copy()
    C = C.copy()
    C[c, c] = 1.0
    G[c, c] = f32(base_z)
    res2 = np.float64(f32(res_f)) ** 2
    for lvl in range(len(level_start) - 1):
        vs = visits[level_start[lvl]:level_start[lvl + 1]]
        x, y = vs[:, 0], vs[:, 1]
        cc = [C[x - 1 + k % 3, y - 1 + k // 3] for k in range(9)]
        gg = [G[x - 1 + k % 3, y - 1 + k // 3] for k in range(9)]
        s = _tree9(cc) + FLT_MIN
        avg = _tree9([a * b for a, b in zip(cc, gg)]) / s
        occ = cc[4]
        G[x, y] = (f32(1.0) - occ) * avg + occ * gg[4]
        fx = (x.astype(np.float32) - f32(c)).astype(np.float64)
        fy = (y.astype(np.float32) - f32(c)).astype(np.float64)
        far = (fx * fx + fy * fy) * res2 > 12.0
        o64 = occ.astype(np.float64)
        dec = np.maximum(o64 - o64 / np.float64(dec_factor), 0.001).astype(np.float32)
        C[x[far], y[far]] = dec[far]
    return G, C


@pytest.mark.parametrize("dim,res,levels", [(33.0, 0.33, None), (99.0, 0.33, 743), (120.0, 0.33, 903)])
def test_spiral_wavefront_schedule_is_exact(dim, res, levels):
    o = Oracle(dim, res)
    n = o.n
    ls, vs = capi.host_spiral_schedule(n)
    c = n // 2 - 1
    assert len(vs) == sum(4 * 2 * (c - p) + 2 for p in range(1, c))       # SURVEY App. C visit count
    if levels is not None:
        assert len(ls) - 1 == levels                                       # SURVEY App. C DAG depth
    # no two visits of a level may touch each other's 3x3 neighbourhood writes
    for lvl in range(0, len(ls) - 1, max(1, (len(ls) - 1) // 40)):
        v = vs[ls[lvl]:ls[lvl + 1]]
        cells = v[:, 0] * n + v[:, 1]
        assert len(np.unique(cells)) == len(cells)
        written = set(cells.tolist())
        for (x, y) in v:
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    if dx or dy:
                        assert (x + dx) * n + (y + dy) not in written
    rng = np.random.default_rng(21)
    o.init_map(0.0, 0.0, 0.0)
    G = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    C = (rng.uniform(0, 1, (n, n)) ** 4).astype(np.float32)
    o.set_layer("ground", G)
    o.set_layer("groundpatch", C)
    o.spiral(0.3)
    Gs, Cs = run_level_schedule(G, C, 0.3, ls, vs, res)
    assert np.array_equal(o.layer("ground"), Gs)
    assert np.array_equal(o.layer("groundpatch"), Cs)


def host_spiral_records(n, res, dist):
    import ctypes as C

    L = capi.load()
    L.gg_host_spiral_records.restype = C.c_int
    L.gg_host_spiral_records.argtypes = [C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    ls, vs = capi.host_spiral_schedule(n)
    recs = np.zeros(4 * len(vs), np.uint32)
    mr = C.c_int(0)
    ok = L.gg_host_spiral_records(n, np.float32(res), dist, recs.ctypes.data_as(C.c_void_p), recs.size, C.byref(mr))
    return ok, mr.value, ls, recs.reshape(-1, 4)


@pytest.mark.parametrize("dist", [1, 2, 3])
@pytest.mark.parametrize("dim,res", [(33.0, 0.33), (99.0, 0.33)])
def test_pipelined_spiral_records_emulation(dim, res, dist):
    """Emulates k_spiral_pipe on the CPU: neighbourhoods are snapshotted `dist` levels before the
    visit, the entries named by the record come from the exchange ring instead, the decay comes
    from the table.  Must reproduce the oracle's sequential sweep bit for bit."""
    o = Oracle(dim, res)
    n = o.n
    ok, max_recent, ls, recs = host_spiral_records(n, res, dist)
    assert ok == 1 and max_recent <= 4
    L = len(ls) - 1
    rng = np.random.default_rng(4)
    o.init_map(0.0, 0.0, 0.0)
    G = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    C = (rng.uniform(0, 1, (n, n)) ** 4).astype(np.float32)
    o.set_layer("ground", G)
    o.set_layer("groundpatch", C)
    o.spiral(0.3)

    c = n // 2 - 1
    G, C = G.copy(), C.copy()
    o64 = C.astype(np.float64)
    D1 = np.maximum(o64 - o64 / 5.0, 0.001).astype(np.float32)
    d64 = D1.astype(np.float64)
    D2 = np.maximum(d64 - d64 / 5.0, 0.001).astype(np.float32)
    C[c, c] = 1.0
    G[c, c] = f32(0.3)
    ring = {}       # level -> (newg, newc) arrays by slot
    snaps = {}      # level -> (cc[9], gg[9]) snapshot taken dist levels ahead

    def snapshot(lvl):
        r = recs[ls[lvl]:ls[lvl + 1]]
        x, y = (r[:, 0] & 0xFFFF).astype(np.int64), (r[:, 0] >> 16).astype(np.int64)
        snaps[lvl] = ([C[x - 1 + q % 3, y - 1 + q // 3].copy() for q in range(9)], [G[x - 1 + q % 3, y - 1 + q // 3].copy() for q in range(9)])

    for lvl in range(min(dist, L)):
        snapshot(lvl)                      # prologue: before level 0 runs
    for lvl in range(L):
        if lvl + dist < L:
            snapshot(lvl + dist)           # prefetch issued at the top of level lvl
        r = recs[ls[lvl]:ls[lvl + 1]]
        x, y = (r[:, 0] & 0xFFFF).astype(np.int64), (r[:, 0] >> 16).astype(np.int64)
        cc, gg = snaps.pop(lvl)
        ents = np.stack([r[:, 1] & 0xFFFF, r[:, 1] >> 16, r[:, 2] & 0xFFFF, r[:, 2] >> 16], axis=1)
        for e in ents.T:
            back = (e >> 14).astype(np.int64)
            q = ((e >> 10) & 15).astype(np.int64)
            slot = (e & 1023).astype(np.int64)
            for b in (1, 2, 3):
                m = back == b
                if not m.any():
                    continue
                assert b <= dist
                src_g, src_c = ring[lvl - b]
                for qq in range(9):
                    mm = m & (q == qq)
                    gg[qq][mm] = src_g[slot[mm]]
                    cc[qq][mm] = src_c[slot[mm]]
        s = _tree9(cc) + FLT_MIN
        avg = _tree9([a * b for a, b in zip(cc, gg)]) / s
        occ = cc[4]
        newg = (f32(1.0) - occ) * avg + occ * gg[4]
        far = (r[:, 3] & 1).astype(bool)
        second = (r[:, 3] & 2).astype(bool)
        newc = np.where(far, np.where(second, D2[x, y], D1[x, y]), occ).astype(np.float32)
        ring[lvl] = (newg.copy(), newc.copy())
        ring.pop(lvl - dist - 1, None)
        G[x, y] = newg
        C[x[far], y[far]] = newc[far]
    assert np.array_equal(o.layer("ground"), G)
    assert np.array_equal(o.layer("groundpatch"), C)


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 16383, 16384, 16385, 40001])
def test_host_cloud_packing(n):
    """gg_filter_cloud_batch repacks PointXYZIR records as x | y | z | ring before the H2D copy."""
    import ctypes as C

    from groundgrid_b200 import synth

    rng = np.random.default_rng(n)
    pts = np.zeros(n, synth.POINT_DTYPE)
    pts.view(np.uint8)[:] = 0xCD          # junk in the padding bytes of the records
    for c in "xyz":
        pts[c] = rng.standard_normal(n).astype(np.float32)
    pts["intensity"] = 7.0
    pts["ring"] = rng.integers(0, 65536, n)
    n_pad = (n + 7) & ~7
    raw = np.full(14 * n_pad + 64, 0xAB, np.uint8)
    off = (-raw.ctypes.data) % 32
    dst = raw[off:off + 14 * n_pad]
    L = capi.load()
    L.gg_host_pack_cloud.restype = C.c_int
    L.gg_host_pack_cloud.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    assert L.gg_host_pack_cloud(pts.ctypes.data_as(C.c_void_p), n, dst.ctypes.data_as(C.c_void_p)) == 0
    f = dst[:12 * n_pad].view(np.float32)
    assert np.array_equal(f[:n], pts["x"]) and np.array_equal(f[n_pad:n_pad + n], pts["y"]) and np.array_equal(f[2 * n_pad:2 * n_pad + n], pts["z"])
    assert np.array_equal(dst[12 * n_pad:].view(np.uint16)[:n], pts["ring"])
    assert np.all(f[n:n_pad] == 0) and np.all(dst[12 * n_pad:].view(np.uint16)[n:] == 0)
    assert np.all(raw[off + 14 * n_pad:] == 0xAB)


def host_spiral_skew(n):
    import ctypes as C

    L = capi.load()
    fn = L.gg_host_spiral_skew
    fn.restype = C.c_int
    fn.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_int]
    hdr = np.zeros(10, np.int32)
    fn(n, hdr.ctypes.data, None, None, None, None, None, None, 0)
    if not hdr[0]:
        return None
    K, KP, rows, levels, row0, lanes, n_irr = (int(v) for v in hdr[1:8])
    pattern = np.zeros(36, np.int32)
    lb, le = np.zeros(lanes, np.int32), np.zeros(lanes, np.int32)
    home = np.zeros(n * n * 4, np.int32)
    ils = np.zeros(levels + 1, np.int32)
    recs = np.zeros(n_irr * 16, np.uint32)
    assert fn(n, hdr.ctypes.data, pattern.ctypes.data, lb.ctypes.data, le.ctypes.data, home.ctypes.data, ils.ctypes.data, recs.ctypes.data, recs.size) == 1
    return dict(K=K, KP=KP, rows=rows, levels=levels, row0=row0, lanes=lanes, pattern=pattern.reshape(4, 9), lane_begin=lb, lane_end=le,
                home=home.reshape(n * n, 4), irr_level_start=ils, irr=recs.reshape(-1, 16))


@pytest.mark.parametrize("dim,res", [(13.2, 0.33), (33.0, 0.33), (99.0, 0.33), (120.0, 0.33), (120.0, 0.2),
                                     (33.33, 0.33), (30.5, 0.5), (81.2, 0.4)])   # the last three: odd cell counts (101, 61, 203)
def test_skewed_spiral_tables_emulation(dim, res):
    """CPU emulation of k_skew -> k_spiral_skew -> k_unskew: lane threads follow the fixed offset
    pattern in (level, ring) space, the irregular warp follows explicit records, neighbourhoods are
    read one level ahead, values written one level ago arrive through the per-lane exchange buffer.
    Must equal the oracle's sequential sweep bit for bit."""
    o = Oracle(dim, res)
    n = o.n
    t = host_spiral_skew(n)
    assert t is not None
    KP, rows, row0, lanes, L = t["KP"], t["rows"], t["row0"], t["lanes"], t["levels"]
    prev_q = [1, 3, 7, 5]
    rng = np.random.default_rng(17)
    o.init_map(0.0, 0.0, 0.0)
    G = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    C = (rng.uniform(0, 1, (n, n)) ** 4).astype(np.float32)
    o.set_layer("ground", G)
    o.set_layer("groundpatch", C)
    o.spiral(0.3)

    # ---- k_skew (column-major cell index = i + j * n)
    c = n // 2 - 1
    Gf, Cf = G.reshape(-1, order="F").copy(), C.reshape(-1, order="F").copy()
    Gf[c + c * n] = f32(0.3)
    Cf[c + c * n] = 1.0
    o64 = Cf.astype(np.float64)
    D1 = np.maximum(o64 - o64 / 5.0, 0.001).astype(np.float32)
    d64 = D1.astype(np.float64)
    D2 = np.maximum(d64 - d64 / 5.0, 0.001).astype(np.float32)
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    fx = (ii.astype(np.float32) - f32(c)).astype(np.float64)
    fy = (jj.astype(np.float32) - f32(c)).astype(np.float64)
    far = ((fx * fx + fy * fy) * np.float64(f32(res)) ** 2 > 12.0).reshape(-1, order="F")
    nslots = 4 * rows * KP
    SKg, SKc, SD = np.zeros(nslots, np.float32), np.zeros(nslots, np.float32), np.full(nslots, -1.0, np.float32)
    home = t["home"]
    for h in range(4):
        m = home[:, h] >= 0
        SKg[home[m, h]] = Gf[m]
        SKc[home[m, h]] = Cf[m]
        if h < 2:
            SD[home[m, h]] = np.where(far[m], (D1 if h == 0 else D2)[m], f32(-1.0))
    # ---- k_spiral_skew
    lane = np.arange(lanes)
    side, kcol = lane // KP, lane % KP
    xg, xc = np.zeros((2, lanes), np.float32), np.zeros((2, lanes), np.float32)
    irr, ils = t["irr"], t["irr_level_start"]

    def reg_lanes(lvl):
        return lane[(t["lane_begin"] <= lvl) & (lvl < t["lane_end"])]

    def fetch(lvl):  # what the prefetch of level `lvl` sees
        rl = reg_lanes(lvl)
        own = (side[rl] * rows + lvl + row0) * KP + kcol[rl]
        idx = own[:, None] + t["pattern"][side[rl]]
        r = irr[ils[lvl]:ils[lvl + 1]]
        iidx = r[:, 1:10].astype(np.int64)
        return (rl, own, SKg[idx].copy(), SKc[idx].copy(), SD[own].copy(), r, SKg[iidx].copy(), SKc[iidx].copy(), SD[r[:, 0].astype(np.int64)].copy())

    def tree(v):
        return ((v[:, 0] + v[:, 1]) + (v[:, 2] + v[:, 3])) + ((v[:, 4] + v[:, 5]) + (v[:, 6] + (v[:, 7] + v[:, 8])))

    def visit(gg, cc, dd):
        s = tree(cc) + FLT_MIN
        avg = tree(cc * gg) / s
        occ = cc[:, 4]
        newg = (f32(1.0) - occ) * avg + occ * gg[:, 4]
        newc = np.where(dd >= 0, dd, occ).astype(np.float32)
        return newg.astype(np.float32), newc

    nxt = fetch(0)
    for lvl in range(L):
        cur = nxt
        if lvl + 1 < L:
            nxt = fetch(lvl + 1)       # issued before this level's stores
        rl, own, gg, cc, dd, r, igg, icc, idd = cur
        pb = (lvl - 1) & 1
        # regular lanes: the previous cell of the lane comes from the exchange buffer
        if len(rl) and lvl > 0:
            pq = np.array(prev_q)[side[rl]]
            gg[np.arange(len(rl)), pq] = xg[pb, rl]
            cc[np.arange(len(rl)), pq] = xc[pb, rl]
        ng, nc = visit(gg, cc, dd)
        # irregular visits: recents name (neighbour, producer lane)
        for w in (10, 11):
            for sh in (0, 16):
                e = (r[:, w] >> sh) & 0xFFFF
                m = e != 0xFFFF
                q = np.where(m, e >> 12, 0).astype(np.int64)
                pl = (e & 4095).astype(np.int64)
                igg[m, q[m]] = xg[pb, pl[m]]
                icc[m, q[m]] = xc[pb, pl[m]]
        ing, inc = visit(igg, icc, idd)
        # stores
        SKg[own], SKc[own] = ng, nc
        xg[lvl & 1, rl], xc[lvl & 1, rl] = ng, nc
        io = r[:, 0].astype(np.int64)
        SKg[io], SKc[io] = ing, inc
        mir = r[:, 12].astype(np.int32)
        mm = mir >= 0
        SKg[mir[mm]], SKc[mir[mm]] = ing[mm], inc[mm]
        il = r[:, 13].astype(np.int64)
        xg[lvl & 1, il], xc[lvl & 1, il] = ing, inc
    # ---- k_unskew
    m = home[:, 0] >= 0
    Gf[m], Cf[m] = SKg[home[m, 0]], SKc[home[m, 0]]
    assert np.array_equal(o.layer("ground"), Gf.reshape(n, n, order="F"))
    assert np.array_equal(o.layer("groundpatch"), Cf.reshape(n, n, order="F"))


def host_spiral_skew_sync(n, M, depth, levels):
    import ctypes as C

    L = capi.load()
    fn = L.gg_host_spiral_skew_sync
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    na = C.c_int(0)
    if not fn(n, M, depth, None, 0, C.byref(na)):
        return None, 0
    req = np.zeros(na.value * levels * 32, np.uint16)
    assert fn(n, M, depth, req.ctypes.data, req.size, C.byref(na)) == 1
    return req.reshape(na.value, levels, 32), na.value


@pytest.mark.parametrize("dim,res,M,policy", [(13.2, 0.33, 32, "random"), (33.0, 0.33, 32, "ahead"), (33.0, 0.33, 64, "random"),
                                              (99.0, 0.33, 64, "ahead"), (99.0, 0.33, 64, "random"), (99.0, 0.33, 160, "ahead"),
                                              (30.5, 0.5, 32, "random"), (120.0, 0.33, 64, "behind"), (81.2, 0.4, 96, "random")])
def test_skewed_spiral_point_to_point_sync_emulation(dim, res, M, policy):
    """The barrier-free variant of k_spiral_skew on the CPU: every agent (a warp of lane threads / the irregular warps)
    runs its levels on its own, held back only by the progress requirements of gg_host_spiral_skew_sync; memory is
    touched exactly when the kernel touches it (slots of level l + 1 loaded during level l, previous-level values from
    the exchange ring of depth 4, stores at the end of the level).  Agents are scheduled adversarially (the one that is
    furthest ahead / behind first, or at random).  Must equal the oracle's sequential sweep bit for bit."""
    DEPTH = 4
    o = Oracle(dim, res)
    n = o.n
    t = host_spiral_skew(n)
    assert t is not None
    KP, rows, row0, lanes, L = t["KP"], t["rows"], t["row0"], t["lanes"], t["levels"]
    M = min(M, KP)
    req, n_agents = host_spiral_skew_sync(n, M, DEPTH, L)
    assert req is not None and n_agents == 4 * M // 32 + 1
    IRR = n_agents - 1
    prev_q = np.array([1, 3, 7, 5])
    rng = np.random.default_rng(23)
    o.init_map(0.0, 0.0, 0.0)
    G = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    C = (rng.uniform(0, 1, (n, n)) ** 4).astype(np.float32)
    o.set_layer("ground", G)
    o.set_layer("groundpatch", C)
    o.spiral(0.3)
    c = n // 2 - 1
    Gf, Cf = G.reshape(-1, order="F").copy(), C.reshape(-1, order="F").copy()
    Gf[c + c * n] = f32(0.3)
    Cf[c + c * n] = 1.0
    o64 = Cf.astype(np.float64)
    D1 = np.maximum(o64 - o64 / 5.0, 0.001).astype(np.float32)
    d64 = D1.astype(np.float64)
    D2 = np.maximum(d64 - d64 / 5.0, 0.001).astype(np.float32)
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    fx = (ii.astype(np.float32) - f32(c)).astype(np.float64)
    fy = (jj.astype(np.float32) - f32(c)).astype(np.float64)
    far = ((fx * fx + fy * fy) * np.float64(f32(res)) ** 2 > 12.0).reshape(-1, order="F")
    nslots = 4 * rows * KP
    SKg, SKc, SD = np.zeros(nslots, np.float32), np.zeros(nslots, np.float32), np.full(nslots, -1.0, np.float32)
    home = t["home"]
    for h in range(4):
        m = home[:, h] >= 0
        SKg[home[m, h]] = Gf[m]
        SKc[home[m, h]] = Cf[m]
        if h < 2:
            SD[home[m, h]] = np.where(far[m], (D1 if h == 0 else D2)[m], f32(-1.0))
    lane = np.arange(lanes)
    side, kcol = lane // KP, lane % KP
    agent_of_lane = (side * M + kcol % M) // 32
    xg, xc = np.zeros((DEPTH, lanes), np.float32), np.zeros((DEPTH, lanes), np.float32)
    irr, ils = t["irr"], t["irr_level_start"]
    lb, le = t["lane_begin"], t["lane_end"]
    lanes_of = [lane[(agent_of_lane == a) & (lb < le)] for a in range(IRR)]

    def tree(v):
        return ((v[:, 0] + v[:, 1]) + (v[:, 2] + v[:, 3])) + ((v[:, 4] + v[:, 5]) + (v[:, 6] + (v[:, 7] + v[:, 8])))

    def visit(gg, cc, dd):
        s = tree(cc) + FLT_MIN
        avg = tree(cc * gg) / s
        occ = cc[:, 4]
        newg = (f32(1.0) - occ) * avg + occ * gg[:, 4]
        return newg.astype(np.float32), np.where(dd >= 0, dd, occ).astype(np.float32)

    def fetch(a, lvl):   # the loads agent a issues for its visits of level lvl
        if lvl >= L:
            return None
        if a == IRR:
            r = irr[ils[lvl]:ils[lvl + 1]]
            if not len(r):
                return None
            iidx = r[:, 1:10].astype(np.int64)
            return (r, SKg[iidx].copy(), SKc[iidx].copy(), SD[r[:, 0].astype(np.int64)].copy())
        la = lanes_of[a]
        rl = la[(lb[la] <= lvl) & (lvl < le[la])]
        if not len(rl):
            return None
        own = (side[rl] * rows + lvl + row0) * KP + kcol[rl]
        idx = own[:, None] + t["pattern"][side[rl]]
        return (rl, own, SKg[idx].copy(), SKc[idx].copy(), SD[own].copy())

    pending = [fetch(a, 0) for a in range(n_agents)]   # prologue of the kernel

    def run(a, lvl):
        cur = pending[a]
        pending[a] = fetch(a, lvl + 1)        # issued before this level's stores
        if cur is None:
            return
        pb = (lvl - 1) % DEPTH
        if a == IRR:
            r, igg, icc, idd = cur
            for w in (10, 11):
                for sh in (0, 16):
                    e = (r[:, w] >> sh) & 0xFFFF
                    m = e != 0xFFFF
                    q = np.where(m, e >> 12, 0).astype(np.int64)
                    pl = (e & 4095).astype(np.int64)
                    igg[m, q[m]] = xg[pb, pl[m]]
                    icc[m, q[m]] = xc[pb, pl[m]]
            ing, inc = visit(igg, icc, idd)
            io = r[:, 0].astype(np.int64)
            SKg[io], SKc[io] = ing, inc
            mir = r[:, 12].astype(np.int32)
            mm = mir >= 0
            SKg[mir[mm]], SKc[mir[mm]] = ing[mm], inc[mm]
            il = r[:, 13].astype(np.int64)
            xg[lvl % DEPTH, il], xc[lvl % DEPTH, il] = ing, inc
        else:
            rl, own, gg, cc, dd = cur
            pq = prev_q[side[rl]]
            gg[np.arange(len(rl)), pq] = xg[pb, rl]
            cc[np.arange(len(rl)), pq] = xc[pb, rl]
            ng, nc = visit(gg, cc, dd)
            SKg[own], SKc[own] = ng, nc
            xg[lvl % DEPTH, rl], xc[lvl % DEPTH, rl] = ng, nc

    progress = np.zeros(32, np.int64)
    progress[n_agents:] = L
    sched = np.random.default_rng(5)
    max_skew = 0
    while np.any(progress[:n_agents] < L):
        runnable = [a for a in range(n_agents) if progress[a] < L and np.all(progress >= req[a, progress[a]])]
        assert runnable, "deadlock"
        if policy == "random":
            a = runnable[sched.integers(len(runnable))]
        elif policy == "ahead":
            a = max(runnable, key=lambda b: progress[b])
        else:
            a = min(runnable, key=lambda b: progress[b])
        run(a, int(progress[a]))
        progress[a] += 1
        act = progress[:n_agents]
        max_skew = max(max_skew, int(act.max() - act.min()))
    m = home[:, 0] >= 0
    Gf[m], Cf[m] = SKg[home[m, 0]], SKc[home[m, 0]]
    assert np.array_equal(o.layer("ground"), Gf.reshape(n, n, order="F"))
    assert np.array_equal(o.layer("groundpatch"), Cf.reshape(n, n, order="F"))
    if policy == "ahead" and n >= 100:
        assert max_skew >= 2   # the point of the exercise: agents really do run apart


@pytest.mark.parametrize("threads,n_jobs,n_points,ring,rounds,lag", [(4, 40, 20000, 8, 4, 3), (3, 17, 50001, 4, 3, 1),
                                                                     (6, 64, 3000, 5, 6, 4), (2, 9, 100, 2, 3, 0),
                                                                     (8, 96, 16385, 32, 5, 6)])
def test_packer_pool_ring_claims_and_cancel(threads, n_jobs, n_points, ring, rounds, lag):
    """The worker pool of gg_filter_cloud_batch without CUDA: ragged clouds through a small staging ring whose slots
    are released `lag` jobs late, jobs taken away from the back (raw path), a cancelled batch, several rounds."""
    import ctypes as C

    L = capi.load()
    f = L.gg_host_packer_selftest
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int]
    for _ in range(3):
        assert f(threads, n_jobs, n_points, ring, rounds, lag) == 0


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (no GPU involved): one JSON line with impl = reference, the metric / unit of the GPU arm, a
    cpu_baseline describing the run and an e2e object repeating the value; it runs the reference's own sources (oracle/_ref)
    whenever that library is available."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--pool", "2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpoints/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    from oracle import ref as refmod

    assert d["cpu_baseline"]["kind"] == ("reference" if refmod.available() else "port")
