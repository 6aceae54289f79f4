"""karto_scanmatcher.create_occupancy_grid on the B200 (reference yag_slam/graph_slam.py:341-342,
ros1/slam_node_ros1:188-209): thin owner of a C-ABI `ysm_occ` handle (include/ysm.h). The pass /
hit counters and the image stay in HBM; `.image` copies the image to the host once, `device_image`
hands the resident image to the ray-walk kernel (`raytracing.raytrace_many(grid, ...)`)."""
import ctypes as C

import numpy as np

from . import _capi


class _Offset(object):
    """grid.offset: .x / .y (+ .yaw = 0 so that ros1 `pose2toPose(grid.offset)` accepts it)."""

    def __init__(self, x, y):
        self.x, self.y, self.yaw = float(x), float(y), 0.0

    def __repr__(self):
        return "Offset(x={}, y={})".format(self.x, self.y)


class OccupancyGrid(object):
    """What create_occupancy_grid returns: .image (uint8 [height][width]: 0 occupied, 200 unknown,
    255 free), .offset (.x, .y), .width, .height (reference ros1/slam_node_ros1:188-209,
    helpers.py:592-603), plus .resolution."""

    def __init__(self, handle, info, device):
        self._h, self._info, self.device = handle, info, int(device)
        self.width, self.height = int(info.width), int(info.height)
        self.offset = _Offset(info.offset_x, info.offset_y)
        self.resolution = float(info.resolution)
        self._image = None

    @property
    def info(self):
        i = self._info
        return dict(rays=int(i.rays), cells_visited=int(i.cells_visited), box_candidates=int(i.box_candidates),
                    cell_fixups=int(i.cell_fixups), launches=int(i.launches))

    @property
    def image(self):
        if self._image is None:
            img = np.empty((self.height, self.width), dtype=np.uint8)
            rc = _capi.lib().ysm_occ_copy_image(self._h, img.ctypes.data)
            if rc != _capi.YSM_OK:
                raise RuntimeError(_capi.lib().ysm_occ_last_error().decode())
            self._image = img
        return self._image

    def counts(self):
        """(pass, hit) uint32 [height][width] -- parity tests."""
        p = np.empty((self.height, self.width), dtype=np.uint32)
        h = np.empty((self.height, self.width), dtype=np.uint32)
        rc = _capi.lib().ysm_occ_copy_counts(self._h, p.ctypes.data, h.ctypes.data)
        if rc != _capi.YSM_OK:
            raise RuntimeError(_capi.lib().ysm_occ_last_error().decode())
        return p, h

    @property
    def device_image(self):
        """Device pointer (int) of the resident [height][width] uint8 image."""
        return int(_capi.lib().ysm_occ_device_image(self._h) or 0)

    def close(self):
        if getattr(self, "_h", None):
            _capi.lib().ysm_occ_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def occupancy_grid_from_arrays(poses, lasers, ranges, beam_ptr, resolution, range_threshold, device=0, stream=0):
    """poses [n][3] sensor poses; lasers [n][4] = min_angle, angular_resolution, min_range, max_range;
    ranges = all raw readings concatenated; beam_ptr [n+1]."""
    poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
    lasers = np.ascontiguousarray(lasers, dtype=np.float64).reshape(-1, 4)
    ranges = np.ascontiguousarray(ranges, dtype=np.float64).reshape(-1)
    beam_ptr = np.ascontiguousarray(beam_ptr, dtype=np.int32).reshape(-1)
    n = len(poses)
    if len(lasers) != n or len(beam_ptr) != n + 1 or (n and beam_ptr[-1] != len(ranges)):
        raise ValueError("occupancy_grid_from_arrays: inconsistent scan arrays")
    s = _capi.YsmOccScans()
    s.n_scans = n
    s.pose, s.laser = poses.ctypes.data, lasers.ctypes.data
    s.ranges, s.beam_ptr = ranges.ctypes.data, beam_ptr.ctypes.data
    s.resolution, s.range_threshold = float(resolution), float(range_threshold)
    L = _capi.lib()
    h = C.c_void_p()
    rc = L.ysm_occ_create(C.byref(s), int(device), C.c_void_p(int(stream)), C.byref(h))
    if rc != _capi.YSM_OK:
        msg = L.ysm_occ_last_error().decode("utf-8", "replace")
        raise {_capi.YSM_EINVAL: ValueError, _capi.YSM_EUNSUP: NotImplementedError}.get(rc, RuntimeError)(msg)
    info = _capi.YsmOccInfo()
    L.ysm_occ_get_info(h, C.byref(info))
    return OccupancyGrid(h, info, device)


def pack_scans(scans):
    """wheel-style LocalizedRangeScan objects (.config, .ranges, .corrected_pose) -> arrays."""
    n = len(scans)
    poses = np.empty((n, 3), np.float64)
    lasers = np.empty((n, 4), np.float64)
    beam_ptr = np.zeros(n + 1, np.int32)
    rr = []
    for i, s in enumerate(scans):
        p, c = s.corrected_pose, s.config
        poses[i] = (p.x, p.y, p.yaw)
        lasers[i] = (c.min_angle, c.angular_resolution, c.min_range, c.max_range)
        r = np.asarray(s.ranges, dtype=np.float64).reshape(-1)
        rr.append(r)
        beam_ptr[i + 1] = beam_ptr[i] + len(r)
    ranges = np.concatenate(rr) if rr else np.zeros(0, np.float64)
    return poses, lasers, ranges, beam_ptr


def create_occupancy_grid(scans, resolution, range_threshold, device=0):
    """create_occupancy_grid(scans, resolution, range_threshold) (reference graph_slam.py:341-342).
    An empty scan list gives None (Karto's CreateFromScans returns NULL)."""
    scans = list(scans)
    if not scans:
        return None
    poses, lasers, ranges, beam_ptr = pack_scans(scans)
    return occupancy_grid_from_arrays(poses, lasers, ranges, beam_ptr, resolution, range_threshold, device=device)
