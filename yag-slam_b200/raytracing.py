"""Mirror of the reference ray-walk API (yag_slam/raytracing.py:63-92) on the CUDA kernel
k_raywalk. `run_raytracing_sweep(img, angles, sx, sy)` returns a list of RayInfo with
.start/.end (Point2 with .x/.y) and .length, like the reference's numba jitclasses;
`raytrace_many` is the batched form (many start cells, one launch) used by the map->graph
splicing caller (yag_slam/splicing.py:87-98)."""
import numpy as np

from . import _capi


class Point2(object):
    def __init__(self, x=0.0, y=0.0):
        self.x, self.y = np.float32(x), np.float32(y)

    @property
    def val(self):
        return (self.x, self.y)


class RayInfo(object):
    def __init__(self, start, end, length):
        self.start, self.end, self._length = start, end, np.float32(length)

    @property
    def length(self):
        return self._length


def raytrace_many(img, angles_deg, starts_xy, device=0, stream=0):
    """(n_starts, n_angles, 5) float32: start.x, start.y, end.x, end.y, length."""
    angles = np.ascontiguousarray(angles_deg, dtype=np.float64)
    starts = np.ascontiguousarray(starts_xy, dtype=np.float64).reshape(-1, 2)
    out = np.zeros((len(starts), len(angles), 5), dtype=np.float32)
    on_dev = hasattr(img, "is_cuda") and bool(img.is_cuda)
    if hasattr(img, "device_image"):  # occupancy.OccupancyGrid: the image is already resident in HBM
        on_dev, h, w, ptr, keep = True, img.height, img.width, img.device_image, img
        device = img.device
    elif on_dev:
        if str(img.dtype) != "torch.uint8" or img.dim() != 2 or not img.is_contiguous():
            raise ValueError("device map must be a contiguous 2-D uint8 tensor")
        if img.device.index is not None and int(img.device.index) != int(device):
            raise ValueError("device map lives on cuda:%d, not on the requested device %d" % (img.device.index, device))
        h, w = int(img.shape[0]), int(img.shape[1])
        ptr, keep = int(img.data_ptr()), img
    else:
        keep = np.ascontiguousarray(img, dtype=np.uint8)
        if keep.ndim != 2:
            raise ValueError("map must be a 2-D uint8 image")
        h, w = keep.shape[:2]
        ptr = keep.ctypes.data
    rc = _capi.lib().ysm_raytrace(ptr, h, w, int(on_dev), angles.ctypes.data, len(angles), starts.ctypes.data,
                                  len(starts), out.ctypes.data, int(device), int(stream))
    del keep
    if rc != _capi.YSM_OK:
        raise RuntimeError(_capi.last_error(None))
    return out


def run_raytracing_sweep(img, angles, sx, sy, device=0):
    res = raytrace_many(img, angles, [(sx, sy)], device=device)[0]
    return [RayInfo(Point2(r[0], r[1]), Point2(r[2], r[3]), r[4]) for r in res]


def trace_ray(img, angle, sx, sy, device=0):
    return run_raytracing_sweep(img, [angle], sx, sy, device=device)[0]
