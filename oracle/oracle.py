"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY).

`KartoOracle` restates karto_scanmatcher.Wrapper.match_scan (reference call site
yag_slam/scan_matching.py:40-42); `raywalk_sweep` restates
yag_slam/raytracing.py:90-92. See the headers of karto_oracle.c / raywalk_oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkarto_oracle.so")

PARAM_FIELDS = [
    "search_size", "resolution", "smear_deviation", "range_threshold",
    "coarse_search_angle_offset", "coarse_angle_resolution", "fine_search_angle_resolution",
    "distance_variance_penalty", "angle_variance_penalty", "minimum_angle_penalty",
    "minimum_distance_penalty",
]


class KoParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in PARAM_FIELDS] + [
        ("use_response_expansion", C.c_int), ("_pad", C.c_int)]


class OcDims(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("offset_x", C.c_double), ("offset_y", C.c_double)]


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("karto_oracle.c", "raywalk_oracle.c", "occgrid_oracle.c", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-s", "-C", _HERE, "clean", "all"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.ko_create.restype = C.c_void_p
        L.ko_create.argtypes = [C.POINTER(KoParams)]
        L.ko_destroy.argtypes = [C.c_void_p]
        L.ko_get_dims.argtypes = [C.c_void_p, ip]
        L.ko_grid_ptr.restype = C.POINTER(C.c_uint8)
        L.ko_grid_ptr.argtypes = [C.c_void_p]
        L.ko_kernel_ptr.restype = C.POINTER(C.c_uint8)
        L.ko_kernel_ptr.argtypes = [C.c_void_p]
        L.ko_lookup_ptr.restype = C.POINTER(C.c_int32)
        L.ko_lookup_ptr.argtypes = [C.c_void_p]
        L.ko_probs_collisions.restype = C.c_long
        L.ko_probs_collisions.argtypes = [C.c_void_p]
        L.ko_point_readings.restype = C.c_int
        L.ko_point_readings.argtypes = [dp, C.c_int] + [C.c_double] * 7 + [dp]
        L.ko_find_valid_points.restype = C.c_int
        L.ko_find_valid_points.argtypes = [dp, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_uint8)]
        L.ko_match.restype = C.c_int
        L.ko_match.argtypes = [C.c_void_p, dp, C.c_int, dp, dp, ip, C.c_int, C.c_int, C.c_int, dp]
        L.ko_match_raw.restype = C.c_int
        L.ko_match_raw.argtypes = [C.c_void_p, dp, C.c_int, C.c_int, dp, dp, ip, C.c_int, C.c_int, C.c_int, dp]
        L.ko_create_map.restype = C.c_void_p
        L.ko_create_map.argtypes = [C.POINTER(KoParams), C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_double,
                                    C.c_double]
        L.ko_match_map.restype = C.c_int
        L.ko_match_map.argtypes = [C.c_void_p, dp, C.c_int, dp, C.c_int, C.c_int, dp]
        L.ko_build_grid.restype = C.c_int
        L.ko_build_grid.argtypes = [C.c_void_p, dp, dp, ip, C.c_int]
        L.ko_compute_offsets.restype = C.c_int
        L.ko_compute_offsets.argtypes = [C.c_void_p, dp, C.c_int, dp, C.c_double, C.c_double, C.c_double]
        L.ko_match_batch.restype = C.c_int
        L.ko_match_batch.argtypes = [C.POINTER(KoParams), C.c_int, dp, ip, ip, ip, dp, ip, ip,
                                     C.c_int, C.c_int, dp, C.c_int]
        L.ko_max_threads.restype = C.c_int
        L.rw_sweep.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, dp, C.c_int, C.c_double,
                               C.c_double, C.POINTER(C.c_float)]
        L.rw_sweep_many.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, dp, C.c_int, dp, C.c_int,
                                    C.POINTER(C.c_float)]
        L.oc_compute_dims.restype = C.c_int
        L.oc_compute_dims.argtypes = [C.c_int, dp, dp, dp, C.POINTER(C.c_int32), C.c_double, C.c_double,
                                      C.POINTER(OcDims)]
        L.oc_render.restype = C.c_int
        L.oc_render.argtypes = [C.c_int, dp, dp, dp, C.POINTER(C.c_int32), C.c_double, C.c_double,
                                C.POINTER(OcDims), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                C.POINTER(C.c_uint8)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


DEFAULTS = dict(
    # yag_slam/helpers.py:339-351 (default_config) + Karto's minimum_distance_penalty
    angle_variance_penalty=0.3, distance_variance_penalty=0.5,
    coarse_search_angle_offset=0.349, coarse_angle_resolution=0.0349,
    fine_search_angle_resolution=0.00349, use_response_expansion=True, range_threshold=20,
    minimum_angle_penalty=0.9, search_size=0.5, resolution=0.01, smear_deviation=0.05,
    minimum_distance_penalty=0.5,
)


def make_params(cfg=None):
    d = dict(DEFAULTS)
    if cfg:
        d.update(cfg)
    p = KoParams()
    for n in PARAM_FIELDS:
        setattr(p, n, float(d[n]))
    p.use_response_expansion = int(bool(d["use_response_expansion"]))
    return p


def point_readings(ranges, min_angle, angular_resolution, min_range, range_threshold, x, y, heading):
    """LocalizedRangeScan::Update: filtered world points, shape (k, 2)."""
    r = np.ascontiguousarray(ranges, dtype=np.float64)
    out = np.empty((len(r), 2), dtype=np.float64)
    k = lib().ko_point_readings(_dp(r), len(r), float(min_angle), float(angular_resolution),
                                float(min_range), float(range_threshold), float(x), float(y),
                                float(heading), _dp(out))
    return out[:k].copy()


def find_valid_points(pts, vpx, vpy):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    mask = np.zeros(len(pts), dtype=np.uint8)
    lib().ko_find_valid_points(_dp(pts), len(pts), float(vpx), float(vpy),
                               mask.ctypes.data_as(C.POINTER(C.c_uint8)))
    return mask


class KartoOracle:
    """CPU restatement of karto_scanmatcher.Wrapper (match_scan on point readings)."""

    def __init__(self, cfg=None):
        self.params = make_params(cfg)
        self._h = lib().ko_create(C.byref(self.params))
        if not self._h:
            raise RuntimeError("oracle: invalid matcher parameters (smear deviation out of bounds?)")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ko_destroy(self._h)
            self._h = None

    def dims(self):
        out = np.zeros(14, dtype=np.int32)
        lib().ko_get_dims(self._h, _ip(out))
        names = ["side", "margin", "roi", "half_kernel", "kernel_size", "border", "width", "height",
                 "stride", "data_size", "n_angles", "n_points", "last_ties", "last_passes"]
        return dict(zip(names, (int(v) for v in out)))

    def kernel(self):
        d = self.dims()
        k = d["kernel_size"]
        return np.ctypeslib.as_array(lib().ko_kernel_ptr(self._h), shape=(k, k)).copy()

    def grid(self):
        d = self.dims()
        return np.ctypeslib.as_array(lib().ko_grid_ptr(self._h), shape=(d["height"], d["stride"])).copy()

    def lookup(self):
        d = self.dims()
        return np.ctypeslib.as_array(lib().ko_lookup_ptr(self._h),
                                     shape=(d["n_angles"], d["n_points"])).copy()

    @staticmethod
    def _pack(base_pts_list):
        counts = np.array([len(b) for b in base_pts_list], dtype=np.int32)
        if len(base_pts_list) and counts.sum() > 0:
            cat = np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.float64).reshape(-1, 2)
                                                       for b in base_pts_list]), dtype=np.float64)
        else:
            cat = np.zeros((1, 2), dtype=np.float64)
        return cat, counts

    def build_grid(self, query_pose, base_pts_list):
        cat, counts = self._pack(base_pts_list)
        pose = np.ascontiguousarray(query_pose, dtype=np.float64)
        lib().ko_build_grid(self._h, _dp(pose), _dp(cat), _ip(counts), len(counts))
        return self.grid()

    def compute_offsets(self, query_pts, pose, angle_center, angle_offset, angle_res):
        q = np.ascontiguousarray(query_pts, dtype=np.float64)
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        lib().ko_compute_offsets(self._h, _dp(q), len(q), _dp(pose), float(angle_center),
                                 float(angle_offset), float(angle_res))
        return self.lookup()

    def match(self, query_pts, query_pose, base_pts_list, do_penalize=True, do_refine=False, n_raw=None):
        """Returns (response, (x, y, heading), cov 3x3). n_raw: RAW range readings of the query (default: as
        many as point readings); beams without any in-range reading raise like Karto."""
        q = np.ascontiguousarray(query_pts, dtype=np.float64).reshape(-1, 2)
        pose = np.ascontiguousarray(query_pose, dtype=np.float64)
        cat, counts = self._pack(base_pts_list)
        out = np.zeros(13, dtype=np.float64)
        qq = q if len(q) else np.zeros((1, 2))
        rc = lib().ko_match_raw(self._h, _dp(qq), len(q), len(q) if n_raw is None else int(n_raw), _dp(pose), _dp(cat),
                                _ip(counts), len(counts), int(do_penalize), int(do_refine), _dp(out))
        if rc != 0:
            raise RuntimeError("Mapper FATAL ERROR - Unable to find best position")
        return float(out[0]), (float(out[1]), float(out[2]), float(out[3])), out[4:].reshape(3, 3).copy()


class KartoMapOracle(KartoOracle):
    """MatchScan against a fixed correlation grid built from a map image (see
    ko_create_map in karto_oracle.c): cells equal to occupied_value are 100 and smeared."""

    def __init__(self, cfg, img, offset_xy, occupied_value=0):
        self.params = make_params(cfg)
        img = np.ascontiguousarray(img, dtype=np.uint8)
        self._h = lib().ko_create_map(C.byref(self.params), img.ctypes.data_as(C.POINTER(C.c_uint8)), img.shape[0],
                                      img.shape[1], int(occupied_value), float(offset_xy[0]), float(offset_xy[1]))
        if not self._h:
            raise RuntimeError("oracle: invalid matcher parameters or empty map")

    def match(self, query_pts, query_pose, do_penalize=True, do_refine=False):
        q = np.ascontiguousarray(query_pts, dtype=np.float64).reshape(-1, 2)
        pose = np.ascontiguousarray(query_pose, dtype=np.float64)
        out = np.zeros(13, dtype=np.float64)
        qq = q if len(q) else np.zeros((1, 2))
        rc = lib().ko_match_map(self._h, _dp(qq), len(q), _dp(pose), int(do_penalize), int(do_refine), _dp(out))
        if rc != 0:
            raise RuntimeError("Mapper FATAL ERROR - Unable to find best position")
        return out

    def match_many(self, pool_xy, scan_start, scan_count, query_scan, query_poses, do_penalize=True, do_refine=False):
        pool_xy = np.asarray(pool_xy, dtype=np.float64).reshape(-1, 2)
        return np.array([self.match(pool_xy[scan_start[q]:scan_start[q] + scan_count[q]], query_poses[i], do_penalize,
                                    do_refine) for i, q in enumerate(query_scan)]).reshape(-1, 13)


def match_batch(cfg, pool_xy, scan_start, scan_count, query_scan, query_poses, base_ptr, base_idx,
                do_penalize=True, do_refine=True, n_threads=0):
    """OpenMP batch driver (one matcher per thread). Returns (n, 13) array."""
    p = make_params(cfg)
    pool_xy = np.ascontiguousarray(pool_xy, dtype=np.float64)
    scan_start = np.ascontiguousarray(scan_start, dtype=np.int32)
    scan_count = np.ascontiguousarray(scan_count, dtype=np.int32)
    query_scan = np.ascontiguousarray(query_scan, dtype=np.int32)
    query_poses = np.ascontiguousarray(query_poses, dtype=np.float64)
    base_ptr = np.ascontiguousarray(base_ptr, dtype=np.int32)
    base_idx = np.ascontiguousarray(base_idx, dtype=np.int32)
    n = len(query_scan)
    out = np.zeros((n, 13), dtype=np.float64)
    rc = lib().ko_match_batch(C.byref(p), n, _dp(pool_xy), _ip(scan_start), _ip(scan_count),
                              _ip(query_scan), _dp(query_poses), _ip(base_ptr), _ip(base_idx),
                              int(do_penalize), int(do_refine), _dp(out), int(n_threads))
    if rc != 0:
        raise RuntimeError("oracle batch failed (rc=%d)" % rc)
    return out


def max_threads():
    return int(lib().ko_max_threads())


def raywalk_sweep(img, angles_deg, sx, sy):
    """run_raytracing_sweep (raytracing.py:90-92): (n, 5) float32 rows start.x,start.y,end.x,end.y,length."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    a = np.ascontiguousarray(angles_deg, dtype=np.float64)
    out = np.zeros((len(a), 5), dtype=np.float32)
    lib().rw_sweep(img.ctypes.data_as(C.POINTER(C.c_uint8)), img.shape[0], img.shape[1], _dp(a), len(a),
                   float(sx), float(sy), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def raywalk_sweep_many(img, angles_deg, starts_xy):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    a = np.ascontiguousarray(angles_deg, dtype=np.float64)
    s = np.ascontiguousarray(starts_xy, dtype=np.float64).reshape(-1, 2)
    out = np.zeros((len(s), len(a), 5), dtype=np.float32)
    lib().rw_sweep_many(img.ctypes.data_as(C.POINTER(C.c_uint8)), img.shape[0], img.shape[1], _dp(a),
                        len(a), _dp(s), len(s), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def occupancy_grid(poses, lasers, ranges, beam_ptr, resolution, range_threshold):
    """Restatement of karto_scanmatcher.create_occupancy_grid (occgrid_oracle.c): returns
    dict(image, passes, hits, width, height, offset_x, offset_y)."""
    poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
    lasers = np.ascontiguousarray(lasers, dtype=np.float64).reshape(-1, 4)
    ranges = np.ascontiguousarray(ranges, dtype=np.float64).reshape(-1)
    beam_ptr = np.ascontiguousarray(beam_ptr, dtype=np.int32).reshape(-1)
    n = len(poses)
    d = OcDims()
    bp = beam_ptr.ctypes.data_as(C.POINTER(C.c_int32))
    if lib().oc_compute_dims(n, _dp(poses), _dp(lasers), _dp(ranges), bp, float(resolution),
                             float(range_threshold), C.byref(d)) != 0:
        return None
    w, h = int(d.width), int(d.height)
    passes = np.zeros((h, w), np.uint32)
    hits = np.zeros((h, w), np.uint32)
    image = np.zeros((h, w), np.uint8)
    lib().oc_render(n, _dp(poses), _dp(lasers), _dp(ranges), bp, float(resolution), float(range_threshold),
                    C.byref(d), passes.ctypes.data_as(C.POINTER(C.c_uint32)),
                    hits.ctypes.data_as(C.POINTER(C.c_uint32)), image.ctypes.data_as(C.POINTER(C.c_uint8)))
    return dict(image=image, passes=passes, hits=hits, width=w, height=h, offset_x=float(d.offset_x),
                offset_y=float(d.offset_y))
