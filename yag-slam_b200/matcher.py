"""ScanMatcherB200: thin Python owner of a C-ABI matcher handle (include/ysm.h).

Host glue only -- all arithmetic of the hot path runs in libysm_b200.so (CUDA kernels +
its libm host runtime). Mirrors the ownership rule of karto_scanmatcher.Wrapper: the matcher
owns its correlation grids / workspaces, scans are borrowed per call.
"""
import ctypes as C

import numpy as np

from . import _capi

DEFAULTS = dict(
    # reference yag_slam/helpers.py:339-351 (default_config); minimum_distance_penalty is
    # Karto's fixed default (not exposed by yag_slam)
    angle_variance_penalty=0.3, distance_variance_penalty=0.5,
    coarse_search_angle_offset=0.349, coarse_angle_resolution=0.0349,
    fine_search_angle_resolution=0.00349, use_response_expansion=True, range_threshold=20,
    minimum_angle_penalty=0.9, search_size=0.5, resolution=0.01, smear_deviation=0.05,
    minimum_distance_penalty=0.5,
)
DEFAULTS_LOOP = dict(DEFAULTS, resolution=0.05, search_size=4.0)  # helpers.py:353-361

_ERRORS = {_capi.YSM_EINVAL: ValueError, _capi.YSM_EUNSUP: NotImplementedError}


def _is_cuda_tensor(x):
    return hasattr(x, "is_cuda") and bool(getattr(x, "is_cuda"))


class ScanMatcherB200(object):
    def __init__(self, cfg=None, device=0, max_slots=0, max_grid_bytes=0, lanes=2, resident_idle_us=0):
        d = dict(DEFAULTS)
        if cfg:
            d.update({k: v for k, v in dict(cfg).items() if k in d})
        self.cfg = d
        p = _capi.YsmParams()
        for n in _capi.PARAM_FIELDS:
            setattr(p, n, float(d[n]))
        p.use_response_expansion = int(bool(d["use_response_expansion"]))
        p.max_slots = int(max_slots)
        p.max_grid_bytes = int(max_grid_bytes)
        p.lanes = int(lanes)
        p.resident_idle_us = int(resident_idle_us)
        self._lib = _capi.lib()
        self._h = C.c_void_p()
        rc = self._lib.ysm_create(C.byref(p), int(device), C.byref(self._h))
        if rc != _capi.YSM_OK:
            msg = _capi.last_error(None)
            self._h = None
            raise _ERRORS.get(rc, RuntimeError)(msg)
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ysm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dims(self):
        out = _capi.YsmDims()
        self._lib.ysm_get_dims(self._h, C.byref(out))
        return {n: int(getattr(out, n)) for n, _ in _capi.YsmDims._fields_}

    def set_debug(self, flags):
        self._lib.ysm_set_debug(self._h, int(flags))

    def launch_count(self):
        return int(self._lib.ysm_launch_count(self._h))

    def last_kernel_ms(self):
        v = [C.c_double() for _ in range(4)]
        self._lib.ysm_last_kernel_ms(self._h, *[C.byref(x) for x in v])
        return dict(zip(("sweep", "build", "reduce", "total"), (x.value for x in v)))

    def last_work(self):
        v = (C.c_int64 * 16)()
        self._lib.ysm_last_work(self._h, v, 16)
        return dict(zip(("lattice_lookups", "sweep_launches", "offset_entries", "poses", "fine_lookups",
                         "base_points", "h2d_bytes", "d2h_bytes", "pruned_sweep_launches", "lookups_issued",
                         "speculative_fine_passes", "lanes", "latency_kernel_launches", "resident_requests",
                         "scan_store_hits", "valid_base_points"),
                        (int(x) for x in v)))

    def ping(self, n=1):
        """Round trips (microseconds) of n empty requests through the resident latency kernel."""
        out = (C.c_double * int(n))()
        rc = self._lib.ysm_debug_ping(self._h, int(n), out)
        if rc != _capi.YSM_OK:
            raise _ERRORS.get(rc, RuntimeError)(_capi.last_error(self._h))
        return np.array(out[:], dtype=np.float64)

    def match_pool(self, pool_xy, scan_start, scan_count, query_scan, query_pose, base_ptr, base_idx,
                   penalty=True, do_fine=False, stream=0, out=None, scan_tag=None, scan_raw_count=None):
        """Batched Wrapper.match_scan over a pool of scans (see ysm_batch in include/ysm.h).

        pool_xy: (n_points, 2) float64 numpy array, or a CUDA tensor already resident in HBM.
        Returns a structured array (dtype _capi.RESULT_DTYPE), one 128-B record per match."""
        scan_start = np.ascontiguousarray(scan_start, dtype=np.int32)
        scan_count = np.ascontiguousarray(scan_count, dtype=np.int32)
        query_scan = np.ascontiguousarray(query_scan, dtype=np.int32)
        query_pose = np.ascontiguousarray(query_pose, dtype=np.float64).reshape(-1, 3)
        base_ptr = np.ascontiguousarray(base_ptr, dtype=np.int32)
        base_idx = np.ascontiguousarray(base_idx, dtype=np.int32)
        n = len(query_scan)
        if len(query_pose) != n or len(base_ptr) != n + 1:
            raise ValueError("query_pose / base_ptr size mismatch")
        b = _capi.YsmBatch()
        if _is_cuda_tensor(pool_xy):
            if str(pool_xy.dtype) != "torch.float64" or not pool_xy.is_contiguous():
                raise ValueError("device pool must be a contiguous float64 tensor")
            b.pool_xy = int(pool_xy.data_ptr())
            b.n_points = int(pool_xy.numel() // 2)
            b.pool_on_device = 1
            keep = pool_xy
        else:
            if hasattr(pool_xy, "numpy") and not isinstance(pool_xy, np.ndarray):
                pool_xy = pool_xy.numpy()  # (pinned) host tensor
            keep = np.ascontiguousarray(pool_xy, dtype=np.float64).reshape(-1, 2)
            b.pool_xy = keep.ctypes.data if len(keep) else None
            b.n_points = len(keep)
            b.pool_on_device = 0
        b.n_matches = n
        b.n_scans = len(scan_start)
        b.scan_start = scan_start.ctypes.data
        b.scan_count = scan_count.ctypes.data
        b.query_scan = query_scan.ctypes.data
        b.query_pose = query_pose.ctypes.data
        b.base_ptr = base_ptr.ctypes.data
        b.base_idx = base_idx.ctypes.data if len(base_idx) else None
        b.do_penalize = int(bool(penalty))
        b.do_refine = int(bool(do_fine))
        if scan_tag is not None:  # content tags (ysm_batch::scan_tag): tagged scans stay resident on the device
            scan_tag = np.ascontiguousarray(scan_tag, dtype=np.uint64)
            if len(scan_tag) != len(scan_start):
                raise ValueError("scan_tag must have one entry per scan")
            b.scan_tag = scan_tag.ctypes.data
        if scan_raw_count is not None:  # raw beams per scan: Karto's "beams but no in-range reading" case raises
            scan_raw_count = np.ascontiguousarray(scan_raw_count, dtype=np.int32)
            if len(scan_raw_count) != len(scan_start):
                raise ValueError("scan_raw_count must have one entry per scan")
            b.scan_raw_count = scan_raw_count.ctypes.data
        if out is None:
            out = np.zeros(n, dtype=_capi.RESULT_DTYPE)
        elif (not isinstance(out, np.ndarray) or out.dtype != _capi.RESULT_DTYPE or out.ndim != 1 or len(out) < n
              or not out.flags["C_CONTIGUOUS"]):
            raise ValueError("out must be a C-contiguous 1-D array of RESULT_DTYPE with room for every match")
        rc = self._lib.ysm_match_batch(self._h, C.byref(b), out.ctypes.data, C.c_void_p(int(stream)))
        del keep
        if rc != _capi.YSM_OK:
            raise _ERRORS.get(rc, RuntimeError)(_capi.last_error(self._h))
        return out

    # ---- parity-test introspection -----------------------------------------------------------
    def debug_grid(self, match=0):
        d = self.dims()
        out = np.zeros((d["height"], d["stride"]), dtype=np.uint8)
        rc = self._lib.ysm_debug_copy_grid(self._h, int(match), out.ctypes.data)
        if rc != _capi.YSM_OK:
            raise RuntimeError(_capi.last_error(self._h))
        return out

    def debug_kernel(self):
        k = self.dims()["kernel_size"]
        out = np.zeros((k, k), dtype=np.uint8)
        rc = self._lib.ysm_debug_copy_kernel(self._h, out.ctypes.data)
        if rc != _capi.YSM_OK:
            raise RuntimeError(_capi.last_error(self._h))
        return out

    def debug_offsets(self, match=0):
        na, npnt = C.c_int32(), C.c_int32()
        rc = self._lib.ysm_debug_copy_offsets(self._h, int(match), None, C.byref(na), C.byref(npnt))
        if rc != _capi.YSM_OK:
            raise RuntimeError(_capi.last_error(self._h))
        out = np.zeros((na.value, npnt.value), dtype=np.int32)
        rc = self._lib.ysm_debug_copy_offsets(self._h, int(match), out.ctypes.data, C.byref(na), C.byref(npnt))
        if rc != _capi.YSM_OK:
            raise RuntimeError(_capi.last_error(self._h))
        return out


class MapMatcherB200(ScanMatcherB200):
    """MatchScan against ONE resident correlation grid built from a map image (SURVEY.md 8(f)-3;
    ysm_create_map in include/ysm.h). Stands in for the reference's unfinished numba map path
    (occupancy_grid_map_to_correlation_grid, yag_slam/helpers.py:24-34, and
    Scan2DMatcherPy.match_scan_sets_with_map, yag_slam/scan_matching.py:124-173) with Karto's grid
    semantics. `img` is uint8 [h][w] (row 0 = minimum y), `offset_xy` the world position of cell
    (0, 0), cfg["resolution"] the map resolution; cells equal to `occupied_value` (0 in the
    reference's map convention, ros1/slam_node_ros1:199-202) are occupied."""

    def __init__(self, cfg, img, offset_xy, occupied_value=0, device=0):
        d = dict(DEFAULTS)
        if cfg:
            d.update({k: v for k, v in dict(cfg).items() if k in d})
        self.cfg = d
        p = _capi.YsmParams()
        for n in _capi.PARAM_FIELDS:
            setattr(p, n, float(d[n]))
        p.use_response_expansion = int(bool(d["use_response_expansion"]))
        img = np.ascontiguousarray(img, dtype=np.uint8)
        if img.ndim != 2 or img.size == 0:
            raise ValueError("map image must be a non-empty 2-D uint8 array")
        self._lib = _capi.lib()
        self._h = C.c_void_p()
        rc = self._lib.ysm_create_map(C.byref(p), img.ctypes.data, img.shape[0], img.shape[1], int(occupied_value),
                                      float(offset_xy[0]), float(offset_xy[1]), int(device), C.byref(self._h))
        if rc != _capi.YSM_OK:
            msg = _capi.last_error(None)
            self._h = None
            raise _ERRORS.get(rc, RuntimeError)(msg)
        self.device = int(device)
        self.map_shape = img.shape

    def correlation_grid(self):
        """The resident grid cropped to the map, uint8 0..100 (what
        occupancy_grid_map_to_correlation_grid returns, scaled by 100)."""
        d = self.dims()
        b = d["border"]
        return self.debug_grid(0)[b:b + self.map_shape[0], b:b + self.map_shape[1]]

    def match_map(self, pool_xy, scan_start, scan_count, query_scan, query_pose, penalty=True, do_fine=False, stream=0):
        """Every query scan (point readings at its initial pose) against the map grid: one record per query."""
        n = len(query_scan)
        return self.match_pool(pool_xy, scan_start, scan_count, query_scan, query_pose, np.zeros(n + 1, np.int32),
                               np.zeros(0, np.int32), penalty, do_fine, stream)


def pack_pool(scans_points):
    """Concatenate per-scan (k_i, 2) point-reading arrays into (pool_xy, scan_start, scan_count)."""
    counts = np.array([len(p) for p in scans_points], dtype=np.int32)
    starts = np.zeros(len(counts), dtype=np.int32)
    if len(counts) > 1:
        starts[1:] = np.cumsum(counts[:-1])
    tot = int(counts.sum())
    pool = np.zeros((tot, 2), dtype=np.float64)
    for s, c, p in zip(starts, counts, scans_points):
        if c:
            pool[s:s + c] = np.asarray(p, dtype=np.float64).reshape(-1, 2)
    return pool, starts, counts
