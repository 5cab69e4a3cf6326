"""ctypes binding of the C-ABI (include/groundgrid_b200.h) -- the same calls the C++ host
classes make.  Python here is plumbing for tests / bench only; there is no Python compute
path and no CPU fallback: every compute call raises GroundGridError without a B200.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build
from .synth import POINT_DTYPE

GG_FLAG_FULL_LAYERS = 1
LABEL_ABSENT, LABEL_GROUND, LABEL_NONGROUND = 0, 49, 99


class GroundGridError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"groundgrid_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    """gg_config == groundgrid::GroundGridConfig (cfg/GroundGrid.cfg:8-21)."""

    _fields_ = [
        ("point_count_cell_variance_threshold", C.c_int),
        ("max_ring", C.c_int),
        ("groundpatch_detection_minimum_threshold", C.c_double),
        ("distance_factor", C.c_double),
        ("minimum_distance_factor", C.c_double),
        ("miminum_point_height_threshold", C.c_double),
        ("minimum_point_height_obstacle_threshold", C.c_double),
        ("outlier_tolerance", C.c_double),
        ("ground_patch_detection_minimum_point_count_threshold", C.c_double),
        ("patch_size_change_distance", C.c_double),
        ("occupied_cells_decrease_factor", C.c_double),
        ("occupied_cells_point_count_factor", C.c_double),
        ("min_outlier_detection_ground_confidence", C.c_double),
        ("thread_count", C.c_int),
    ]


class ScanDesc(C.Structure):
    _fields_ = [
        ("slot", C.c_int),
        ("_reserved", C.c_int),
        ("n_points", C.c_size_t),
        ("origin", C.c_float * 3),
        ("_pad", C.c_float),
        ("base_z", C.c_double),
    ]


_lib = None


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Loads libgroundgrid_b200.so (building it in-tree with nvcc if it is missing/stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        try:
            _build.build_core()
        except Exception:
            if not os.path.exists(_build.LIB):
                raise
    L = C.CDLL(_build.LIB)
    vp, i, d, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
    sig = {
        "gg_default_config": (None, [C.POINTER(Config)]),
        "gg_last_error": (C.c_char_p, []),
        "gg_create": (i, [d, C.c_float, i, i, sz, C.c_uint, vp, C.POINTER(vp)]),
        "gg_destroy": (i, [vp]),
        "gg_cells_per_side": (i, [vp]),
        "gg_num_slots": (i, [vp]),
        "gg_set_config": (i, [vp, C.POINTER(Config)]),
        "gg_get_config": (i, [vp, C.POINTER(Config)]),
        "gg_init_map": (i, [vp, i, d, d, d]),
        "gg_update_pose": (i, [vp, i, d, d, vp, C.POINTER(i)]),
        "gg_update_pose_batch": (i, [vp, i, vp, vp, vp, vp]),
        "gg_get_map_position": (i, [vp, i, vp]),
        "gg_set_map_position": (i, [vp, i, d, d]),
        "gg_filter_cloud": (i, [vp, i, vp, sz, vp, d, vp, vp, vp, C.POINTER(sz)]),
        "gg_filter_cloud_batch": (i, [vp, i, vp, vp, vp]),
        "gg_filter_cloud_batch_begin": (i, [vp, i, vp, vp, vp, C.POINTER(i)]),
        "gg_filter_cloud_batch_wait": (i, [vp, i]),
        "gg_upload_points": (i, [vp, i, vp, sz]),
        "gg_run_scans": (i, [vp, i, vp, i]),
        "gg_download_labels": (i, [vp, i, vp, sz]),
        "gg_synchronize": (i, [vp]),
        "gg_run_scans_device": (i, [vp, i, vp, vp, i]),
        "gg_upload_cloud_msg": (i, [vp, i, vp, sz, i, vp, vp]),
        "gg_terrain_image": (i, [vp, i, vp]),
        "gg_layer_image_u8": (i, [vp, i, C.c_char_p, vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "gg_get_point_classes": (i, [vp, i, vp, sz]),
        "gg_detect_ground_patches": (i, [vp, i]),
        "gg_detect_ground_patch": (i, [vp, i, i, i, i]),
        "gg_spiral_ground_interpolation": (i, [vp, i, d]),
        "gg_interpolate_cell": (i, [vp, i, i, i]),
        "gg_eval_accumulate": (i, [vp, i]),
        "gg_eval_read": (i, [vp, vp, i]),
        "gg_profile_enable": (i, [vp, i]),
        "gg_profile_read": (i, [vp, vp, vp, i]),
        "gg_profile_kernel_count": (i, []),
        "gg_profile_kernel_name": (C.c_char_p, [i]),
        "gg_get_output": (i, [vp, i, vp, vp, C.POINTER(sz)]),
        "gg_get_layer": (i, [vp, i, C.c_char_p, vp]),
        "gg_set_layer": (i, [vp, i, C.c_char_p, vp]),
        "gg_layer_device_ptr": (i, [vp, i, C.c_char_p, C.POINTER(vp)]),
        "gg_stream": (vp, [vp]),
        "gg_num_streams": (i, [vp]),
        "gg_host_pack_threads": (i, [vp]),
        "gg_last_batch_transfer": (i, [vp, C.POINTER(C.c_size_t)]),
        "gg_fork_streams": (i, [vp]),
        "gg_join_streams": (i, [vp]),
        "gg_kernel_launches": (C.c_uint64, [vp]),
        "gg_spiral_schedule_info": (i, [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
        # host-only helpers (no device needed)
        "gg_host_cells_per_side": (i, [d, C.c_float]),
        "gg_host_expected_points": (i, [d, C.c_float, vp]),
        "gg_host_spiral_schedule": (i, [i, vp, i, vp, i, C.POINTER(i), C.POINTER(i)]),
        "gg_host_move_map": (i, [d, vp, d, d, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _check(rc):
    if rc != 0:
        raise GroundGridError(rc, load().gg_last_error().decode(errors="replace"))


# ---- host-only helpers --------------------------------------------------------------------
def host_cells_per_side(dimension_m, resolution):
    return load().gg_host_cells_per_side(float(dimension_m), np.float32(resolution))


def host_expected_points(dimension_m, resolution):
    n = host_cells_per_side(dimension_m, resolution)
    out = np.empty((n, n), np.float32, order="F")
    load().gg_host_expected_points(float(dimension_m), np.float32(resolution), _ptr(out))
    return out


def host_spiral_schedule(n):
    """(level_start[int32, levels+1], visits[(x, y) per visit, grouped by level])."""
    L = load()
    nl, nv = C.c_int(0), C.c_int(0)
    L.gg_host_spiral_schedule(n, None, 0, None, 0, C.byref(nl), C.byref(nv))
    ls = np.zeros(nl.value + 1, np.int32)
    vs = np.zeros(nv.value, np.uint32)
    L.gg_host_spiral_schedule(n, _ptr(ls), ls.size, _ptr(vs), vs.size, C.byref(nl), C.byref(nv))
    return ls, np.stack([vs & 0xFFFF, vs >> 16], axis=1).astype(np.int32)


def host_move_map(res, pos_xy, new_xy):
    pos = np.array(pos_xy, np.float64)
    shift = np.zeros(2, np.int32)
    moved = load().gg_host_move_map(float(res), _ptr(pos), float(new_xy[0]), float(new_xy[1]), _ptr(shift))
    return bool(moved), pos, (int(shift[0]), int(shift[1]))


# ---- device handle ------------------------------------------------------------------------
class GroundGridB200:
    """One handle = `n_slots` independent GroundGrid maps on one GPU (see the C header)."""

    def __init__(self, dimension_m=120.0, resolution=0.33, device=0, n_slots=1, max_points=131072,
                 full_layers=False, stream=None):
        self._l = load()
        h = C.c_void_p()
        flags = GG_FLAG_FULL_LAYERS if full_layers else 0
        _check(self._l.gg_create(float(dimension_m), np.float32(resolution), int(device), int(n_slots), int(max_points),
                                 flags, C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self.device = int(device)
        self.n = self._l.gg_cells_per_side(h)
        self.n_slots = n_slots
        self.max_points = max_points

    def close(self):
        if getattr(self, "_h", None):
            self._l.gg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- config / state
    def set_config(self, **kw):
        cfg = Config()
        _check(self._l.gg_get_config(self._h, C.byref(cfg)))
        for k, v in kw.items():
            if not hasattr(cfg, k):
                raise KeyError(k)
            setattr(cfg, k, v)
        _check(self._l.gg_set_config(self._h, C.byref(cfg)))

    def init_map(self, x, y, z, slot=0):
        _check(self._l.gg_init_map(self._h, slot, x, y, z))

    def update_pose(self, x, y, T_base_from_map, slot=0):
        T = np.ascontiguousarray(T_base_from_map, dtype=np.float64).reshape(12)
        moved = C.c_int(0)
        _check(self._l.gg_update_pose(self._h, slot, x, y, _ptr(T), C.byref(moved)))
        return bool(moved.value)

    def update_pose_batch(self, slots, xy, T):
        slots = np.ascontiguousarray(slots, np.int32)
        xy = np.ascontiguousarray(xy, np.float64).reshape(len(slots), 2)
        T = np.ascontiguousarray(T, np.float64).reshape(len(slots), 12)
        moved = np.zeros(len(slots), np.int32)
        _check(self._l.gg_update_pose_batch(self._h, len(slots), _ptr(slots), _ptr(xy), _ptr(T), _ptr(moved)))
        return moved.astype(bool)

    def position(self, slot=0):
        xy = np.zeros(2, np.float64)
        _check(self._l.gg_get_map_position(self._h, slot, _ptr(xy)))
        return xy

    def set_position(self, x, y, slot=0):
        _check(self._l.gg_set_map_position(self._h, slot, x, y))

    def layer(self, name, slot=0):
        out = np.empty((self.n, self.n), np.float32, order="F")
        _check(self._l.gg_get_layer(self._h, slot, name.encode(), _ptr(out)))
        return out

    def set_layer(self, name, arr, slot=0):
        a = np.asfortranarray(arr, dtype=np.float32)
        assert a.shape == (self.n, self.n)
        _check(self._l.gg_set_layer(self._h, slot, name.encode(), _ptr(a)))

    def layer_device_ptr(self, name, slot=0):
        p = C.c_void_p()
        _check(self._l.gg_layer_device_ptr(self._h, slot, name.encode(), C.byref(p)))
        return p.value

    @property
    def stream(self):
        return self._l.gg_stream(self._h)

    @property
    def n_streams(self):
        return self._l.gg_num_streams(self._h)

    @property
    def host_pack_threads(self):
        return self._l.gg_host_pack_threads(self._h)

    def last_batch_transfer(self):
        """(scans packed, scans raw, packed H2D bytes, raw H2D bytes, feed us, total us, packers' pack us,
        packers' slot-wait us, feeder idle us) of the last batch call."""
        info = (C.c_size_t * 9)()
        _check(self._l.gg_last_batch_transfer(self._h, info))
        return tuple(int(v) for v in info)

    def fork_streams(self):
        _check(self._l.gg_fork_streams(self._h))

    def join_streams(self):
        _check(self._l.gg_join_streams(self._h))

    @property
    def kernel_launches(self):
        return int(self._l.gg_kernel_launches(self._h))

    def spiral_schedule_info(self):
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        _check(self._l.gg_spiral_schedule_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    # -- the reference-facing call: GroundSegmentation::filter_cloud with host buffers
    def filter_cloud(self, points, origin, base_z, slot=0, want_index=False, want_cloud=False):
        pts = np.ascontiguousarray(points, dtype=POINT_DTYPE)
        n = pts.shape[0]
        labels = np.zeros(n, np.uint8)
        org = np.ascontiguousarray(origin, dtype=np.float32)
        index = np.zeros(n, np.uint32) if want_index else None
        cloud = np.zeros(n, POINT_DTYPE) if want_cloud else None
        nout = C.c_size_t(0)
        _check(self._l.gg_filter_cloud(self._h, slot, _ptr(pts), n, _ptr(org), float(base_z), _ptr(labels), _ptr(index),
                                       _ptr(cloud), C.byref(nout) if (want_index or want_cloud) else None))
        if want_index or want_cloud:
            k = nout.value
            return labels, (index[:k] if want_index else None), (cloud[:k] if want_cloud else None)
        return labels

    # -- device-resident pieces
    def upload_points(self, points, slot=0):
        pts = np.ascontiguousarray(points, dtype=POINT_DTYPE)
        _check(self._l.gg_upload_points(self._h, slot, _ptr(pts), pts.shape[0]))
        return pts  # keep alive until synchronize()

    def upload_points_ptr(self, ptr, n, slot=0):
        _check(self._l.gg_upload_points(self._h, slot, C.c_void_p(ptr), n))

    @staticmethod
    def make_descs(slots, n_points, origins, base_z):
        arr = (ScanDesc * len(slots))()
        for k, s in enumerate(slots):
            arr[k].slot = int(s)
            arr[k].n_points = int(n_points[k])
            arr[k].origin[0], arr[k].origin[1], arr[k].origin[2] = [float(v) for v in origins[k]]
            arr[k].base_z = float(base_z[k])
        return arr

    def run_scans(self, descs, stop_after=0):
        _check(self._l.gg_run_scans(self._h, len(descs), descs, stop_after))

    def run_scans_device(self, descs, dev_ptrs, stop_after=0):
        """dev_ptrs: device addresses (ints) of the per-scan clouds (32-byte records)."""
        pp = (C.c_void_p * len(descs))(*dev_ptrs)
        _check(self._l.gg_run_scans_device(self._h, len(descs), descs, pp, stop_after))

    # -- steps next to the path
    def upload_cloud_msg(self, raw, n_points, point_step, field_offsets, T_map_from_frame=None, slot=0):
        """raw: uint8 array of the PointCloud2 payload; field_offsets: x, y, z, intensity, ring (-1 absent)."""
        raw = np.ascontiguousarray(raw, np.uint8)
        off = np.ascontiguousarray(field_offsets, np.int32)
        T = None if T_map_from_frame is None else np.ascontiguousarray(T_map_from_frame, np.float64).reshape(12)
        _check(self._l.gg_upload_cloud_msg(self._h, slot, _ptr(raw), int(n_points), int(point_step), _ptr(off), _ptr(T)))
        return raw

    def terrain_image(self, slot=0):
        img = np.zeros((self.n, self.n, 3), np.float32)
        _check(self._l.gg_terrain_image(self._h, slot, _ptr(img)))
        return img

    def layer_image_u8(self, name, slot=0):
        """(N x N uint8 image indexed [i, j], lower, upper): what toImage<unsigned char, 1> hands to cv::applyColorMap."""
        img = np.zeros((self.n, self.n), np.uint8)
        lo, hi = C.c_float(0), C.c_float(0)
        _check(self._l.gg_layer_image_u8(self._h, slot, name.encode(), _ptr(img), C.byref(lo), C.byref(hi)))
        return img, lo.value, hi.value

    # -- the reference's per-phase methods
    def point_classes(self, n, slot=0):
        codes = np.zeros(n, np.uint32)
        _check(self._l.gg_get_point_classes(self._h, slot, _ptr(codes), n))
        return codes

    def detect_ground_patches(self, slot=0):
        _check(self._l.gg_detect_ground_patches(self._h, slot))

    def detect_ground_patch(self, size, i, j, slot=0):
        _check(self._l.gg_detect_ground_patch(self._h, slot, size, i, j))

    def spiral_ground_interpolation(self, base_z, slot=0):
        _check(self._l.gg_spiral_ground_interpolation(self._h, slot, float(base_z)))

    def interpolate_cell(self, x, y, slot=0):
        _check(self._l.gg_interpolate_cell(self._h, slot, x, y))

    def eval_accumulate(self, slot=0):
        _check(self._l.gg_eval_accumulate(self._h, slot))

    def eval_read(self, reset=False):
        counts = np.zeros((1024, 2), np.uint64)
        _check(self._l.gg_eval_read(self._h, _ptr(counts), 1 if reset else 0))
        return counts

    def profile_enable(self, on=True):
        _check(self._l.gg_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset=True):
        """{kernel name: (total ms, launches)} measured with CUDA events on the launching stream."""
        k = self._l.gg_profile_kernel_count()
        ms = np.zeros(k, np.float64)
        cnt = np.zeros(k, np.uint32)
        _check(self._l.gg_profile_read(self._h, _ptr(ms), _ptr(cnt), 1 if reset else 0))
        return {self._l.gg_profile_kernel_name(j).decode(): (float(ms[j]), int(cnt[j])) for j in range(k) if cnt[j]}

    def download_labels(self, n, slot=0, out=None):
        out = np.zeros(n, np.uint8) if out is None else out
        _check(self._l.gg_download_labels(self._h, slot, _ptr(out), n))
        return out

    def download_labels_ptr(self, ptr, n, slot=0):
        _check(self._l.gg_download_labels(self._h, slot, C.c_void_p(ptr), n))

    def filter_cloud_batch_ptrs(self, descs, point_ptrs, label_ptrs):
        """Host pointers (ints) per scan; see gg_filter_cloud_batch."""
        n = len(descs)
        pp = (C.c_void_p * n)(*point_ptrs)
        lp = (C.c_void_p * n)(*label_ptrs) if label_ptrs is not None else None
        _check(self._l.gg_filter_cloud_batch(self._h, n, descs, pp, lp))

    def filter_cloud_batch_begin(self, descs, point_ptrs, label_ptrs):
        """First half of filter_cloud_batch_ptrs: returns a ticket once everything is enqueued."""
        n = len(descs)
        pp = (C.c_void_p * n)(*point_ptrs)
        lp = (C.c_void_p * n)(*label_ptrs) if label_ptrs is not None else None
        ticket = C.c_int(-1)
        _check(self._l.gg_filter_cloud_batch_begin(self._h, n, descs, pp, lp, C.byref(ticket)))
        return ticket.value

    def filter_cloud_batch_wait(self, ticket):
        _check(self._l.gg_filter_cloud_batch_wait(self._h, ticket))

    def synchronize(self):
        _check(self._l.gg_synchronize(self._h))

    def get_output(self, slot=0, want_cloud=False):
        n = self.max_points
        index = np.zeros(n, np.uint32)
        cloud = np.zeros(n, POINT_DTYPE) if want_cloud else None
        nout = C.c_size_t(0)
        _check(self._l.gg_get_output(self._h, slot, _ptr(index), _ptr(cloud), C.byref(nout)))
        k = nout.value
        return index[:k], (cloud[:k] if want_cloud else None)

    def run_single(self, points, origin, base_z, slot=0, stop_after=0):
        """Upload + run (optionally only the first phases) + sync; returns labels if the run was complete."""
        keep = self.upload_points(points, slot)
        d = self.make_descs([slot], [keep.shape[0]], [origin], [base_z])
        self.run_scans(d, stop_after)
        labels = self.download_labels(keep.shape[0], slot) if stop_after == 0 else None
        self.synchronize()
        return labels
