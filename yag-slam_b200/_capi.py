"""ctypes binding of the C ABI in include/ysm.h (libysm_b200.so). No CPU fallback: loading
fails loudly if the CUDA library has not been built."""
import ctypes as C
import os

import numpy as np

from .build import SO_PATH

PARAM_FIELDS = [
    "search_size", "resolution", "smear_deviation", "range_threshold",
    "coarse_search_angle_offset", "coarse_angle_resolution", "fine_search_angle_resolution",
    "distance_variance_penalty", "angle_variance_penalty", "minimum_angle_penalty",
    "minimum_distance_penalty",
]

YSM_OK, YSM_EINVAL, YSM_ECUDA, YSM_ENOMEM, YSM_EMATCH, YSM_EUNSUP = 0, -1, -2, -3, -4, -5
DEBUG_KEEP_GRIDS, DEBUG_TIME_KERNELS, DEBUG_NO_PRUNE, DEBUG_NO_SPECULATE, DEBUG_NO_MEGA, DEBUG_NO_CANDLISTS = 1, 2, 4, 8, 16, 32
DEBUG_NO_HALF_LISTS = 64


class YsmParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in PARAM_FIELDS] + [
        ("use_response_expansion", C.c_int32), ("max_slots", C.c_int32), ("max_grid_bytes", C.c_int64),
        ("lanes", C.c_int32), ("resident_idle_us", C.c_int32)]


class YsmDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("side", "margin", "roi", "half_kernel", "kernel_size", "border",
                                         "width", "height", "stride", "slots")] + [("grid_bytes", C.c_int64)]


class YsmBatch(C.Structure):
    _fields_ = [
        ("n_matches", C.c_int32), ("n_scans", C.c_int32), ("n_points", C.c_int64),
        ("pool_xy", C.c_void_p), ("scan_start", C.c_void_p), ("scan_count", C.c_void_p),
        ("query_scan", C.c_void_p), ("query_pose", C.c_void_p), ("base_ptr", C.c_void_p),
        ("base_idx", C.c_void_p), ("do_penalize", C.c_int32), ("do_refine", C.c_int32),
        ("pool_on_device", C.c_int32), ("_pad", C.c_int32), ("scan_tag", C.c_void_p),
        ("scan_raw_count", C.c_void_p)]


class YsmOccScans(C.Structure):
    _fields_ = [("n_scans", C.c_int32), ("_pad", C.c_int32), ("pose", C.c_void_p), ("laser", C.c_void_p),
                ("ranges", C.c_void_p), ("beam_ptr", C.c_void_p), ("resolution", C.c_double),
                ("range_threshold", C.c_double)]


class YsmOccInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("offset_x", C.c_double), ("offset_y", C.c_double),
                ("resolution", C.c_double), ("rays", C.c_int64), ("cells_visited", C.c_int64),
                ("box_candidates", C.c_int32), ("cell_fixups", C.c_int32), ("launches", C.c_int32),
                ("_pad", C.c_int32)]


class YsmChainQuery(C.Structure):
    _fields_ = [("n_vertices", C.c_int32), ("n_queries", C.c_int32), ("pose_xy", C.c_void_p), ("hash_xy", C.c_void_p),
                ("adj_ptr", C.c_void_p), ("adj_idx", C.c_void_p), ("query_vertex", C.c_void_p),
                ("loop_search_dist", C.c_double), ("crude_r2", C.c_double), ("near_dist_sq", C.c_double),
                ("min_chain_size", C.c_int32), ("_pad", C.c_int32)]


RESULT_DTYPE = np.dtype([("response", "<f8"), ("x", "<f8"), ("y", "<f8"), ("heading", "<f8"),
                         ("cov", "<f8", (9,)), ("n_passes", "<i4"), ("n_ties", "<i4"),
                         ("status", "<i4"), ("_pad", "<i4"), ("_reserved", "<f8")])
assert RESULT_DTYPE.itemsize == 128

EXPORTS = [
    "ysm_create", "ysm_create_map", "ysm_destroy", "ysm_last_error", "ysm_get_dims", "ysm_match_batch",
    "ysm_point_readings", "ysm_point_readings_batch", "ysm_raytrace", "ysm_set_debug", "ysm_debug_copy_grid",
    "ysm_debug_copy_kernel", "ysm_debug_copy_offsets", "ysm_launch_count", "ysm_last_kernel_ms", "ysm_last_work",
    "ysm_debug_ping",
    "ysm_occ_create", "ysm_occ_destroy", "ysm_occ_get_info", "ysm_occ_copy_image", "ysm_occ_copy_counts",
    "ysm_occ_device_image", "ysm_occ_last_error",
    "ysm_chains_find", "ysm_chains_get_counts", "ysm_chains_copy", "ysm_chains_destroy", "ysm_chains_last_error",
]

_lib = None


def lib():
    """Loads libysm_b200.so (raises if it is missing -- there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            "libysm_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback." % SO_PATH)
    L = C.CDLL(SO_PATH)
    vp, i32, f64 = C.c_void_p, C.c_int32, C.c_double
    L.ysm_create.restype = C.c_int
    L.ysm_create.argtypes = [C.POINTER(YsmParams), C.c_int, C.POINTER(vp)]
    L.ysm_create_map.restype = C.c_int
    L.ysm_create_map.argtypes = [C.POINTER(YsmParams), vp, i32, i32, i32, f64, f64, C.c_int, C.POINTER(vp)]
    L.ysm_destroy.restype = None
    L.ysm_destroy.argtypes = [vp]
    L.ysm_last_error.restype = C.c_char_p
    L.ysm_last_error.argtypes = [vp]
    L.ysm_get_dims.restype = C.c_int
    L.ysm_get_dims.argtypes = [vp, C.POINTER(YsmDims)]
    L.ysm_match_batch.restype = C.c_int
    L.ysm_match_batch.argtypes = [vp, C.POINTER(YsmBatch), vp, vp]
    L.ysm_point_readings.restype = C.c_int
    L.ysm_point_readings.argtypes = [vp, i32, f64, f64, f64, f64, f64, f64, f64, vp, C.POINTER(i32)]
    L.ysm_point_readings_batch.restype = C.c_int
    L.ysm_point_readings_batch.argtypes = [vp, vp, i32, vp, vp, i32, f64, f64, f64, f64, vp, vp, vp, C.POINTER(C.c_int64)]
    L.ysm_raytrace.restype = C.c_int
    L.ysm_raytrace.argtypes = [vp, i32, i32, i32, vp, i32, vp, i32, vp, C.c_int, vp]
    L.ysm_set_debug.restype = C.c_int
    L.ysm_set_debug.argtypes = [vp, i32]
    L.ysm_debug_copy_grid.restype = C.c_int
    L.ysm_debug_copy_grid.argtypes = [vp, i32, vp]
    L.ysm_debug_copy_kernel.restype = C.c_int
    L.ysm_debug_copy_kernel.argtypes = [vp, vp]
    L.ysm_debug_copy_offsets.restype = C.c_int
    L.ysm_debug_copy_offsets.argtypes = [vp, i32, vp, C.POINTER(i32), C.POINTER(i32)]
    L.ysm_launch_count.restype = C.c_int64
    L.ysm_launch_count.argtypes = [vp]
    L.ysm_last_kernel_ms.restype = C.c_int
    L.ysm_last_kernel_ms.argtypes = [vp] + [C.POINTER(f64)] * 4
    L.ysm_last_work.restype = C.c_int
    L.ysm_last_work.argtypes = [vp, C.POINTER(C.c_int64), i32]
    L.ysm_debug_ping.restype = C.c_int
    L.ysm_debug_ping.argtypes = [vp, i32, C.POINTER(f64)]
    L.ysm_occ_create.restype = C.c_int
    L.ysm_occ_create.argtypes = [C.POINTER(YsmOccScans), C.c_int, vp, C.POINTER(vp)]
    L.ysm_occ_destroy.restype = None
    L.ysm_occ_destroy.argtypes = [vp]
    L.ysm_occ_get_info.restype = C.c_int
    L.ysm_occ_get_info.argtypes = [vp, C.POINTER(YsmOccInfo)]
    L.ysm_occ_copy_image.restype = C.c_int
    L.ysm_occ_copy_image.argtypes = [vp, vp]
    L.ysm_occ_copy_counts.restype = C.c_int
    L.ysm_occ_copy_counts.argtypes = [vp, vp, vp]
    L.ysm_occ_device_image.restype = vp
    L.ysm_occ_device_image.argtypes = [vp]
    L.ysm_occ_last_error.restype = C.c_char_p
    L.ysm_occ_last_error.argtypes = []
    L.ysm_chains_find.restype = C.c_int
    L.ysm_chains_find.argtypes = [C.POINTER(YsmChainQuery), C.c_int, vp, C.POINTER(vp)]
    L.ysm_chains_get_counts.restype = C.c_int
    L.ysm_chains_get_counts.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(f64)]
    L.ysm_chains_copy.restype = C.c_int
    L.ysm_chains_copy.argtypes = [vp, vp, vp, vp]
    L.ysm_chains_destroy.restype = None
    L.ysm_chains_destroy.argtypes = [vp]
    L.ysm_chains_last_error.restype = C.c_char_p
    L.ysm_chains_last_error.argtypes = []
    _lib = L
    return L


def last_error(handle=None):
    s = lib().ysm_last_error(handle)
    return s.decode("utf-8", "replace") if s else ""


def point_readings(ranges, min_angle, angular_resolution, min_range, range_threshold, x, y, heading):
    """LocalizedRangeScan::Update (host libm inside the library): (k, 2) float64 world points."""
    r = np.ascontiguousarray(ranges, dtype=np.float64)
    out = np.empty((max(1, len(r)), 2), dtype=np.float64)
    n = C.c_int32(0)
    rc = lib().ysm_point_readings(r.ctypes.data, len(r), float(min_angle), float(angular_resolution),
                                  float(min_range), float(range_threshold), float(x), float(y),
                                  float(heading), out.ctypes.data, C.byref(n))
    if rc != YSM_OK:
        raise RuntimeError("ysm_point_readings failed (%d)" % rc)
    return out[:n.value].copy()


def point_readings_batch(ranges, beam_ptr, src, pose, min_angle, angular_resolution, min_range, range_threshold):
    """LocalizedRangeScan::Update for a batch of scans in one call: scan i = source scan src[i] at pose[i].
    Returns (pool_xy (n_points, 2), starts, counts)."""
    ranges = np.ascontiguousarray(ranges, dtype=np.float64)
    beam_ptr = np.ascontiguousarray(beam_ptr, dtype=np.int32)
    src = np.ascontiguousarray(src, dtype=np.int32)
    pose = np.ascontiguousarray(pose, dtype=np.float64).reshape(-1, 3)
    n = len(src)
    if len(pose) != n:
        raise ValueError("one pose per scan")
    cap = int((beam_ptr[1:] - beam_ptr[:-1])[src].sum()) if n else 0
    out = np.empty((max(cap, 1), 2), dtype=np.float64)
    starts, counts = np.zeros(n, np.int32), np.zeros(n, np.int32)
    tot = C.c_int64(0)
    rc = lib().ysm_point_readings_batch(ranges.ctypes.data, beam_ptr.ctypes.data, len(beam_ptr) - 1, src.ctypes.data,
                                        pose.ctypes.data, n, float(min_angle), float(angular_resolution),
                                        float(min_range), float(range_threshold), out.ctypes.data, starts.ctypes.data,
                                        counts.ctypes.data, C.byref(tot))
    if rc != YSM_OK:
        raise RuntimeError("ysm_point_readings_batch failed (%d)" % rc)
    return out[:tot.value], starts, counts
