"""Generates tests/golden/mapgrid_golden.npz by running the REFERENCE's numba
occupancy_grid_map_to_correlation_grid (/root/reference/yag_slam/helpers.py:24-34, unmodified) in the
build container, for parameter pairs where its kernel has Karto's size (4*round(s/r)+1 == 2*Round(2s/r)+1,
i.e. smear_deviation / resolution integral). Karto quantises the kernel to Round(100*z) bytes; the
vectors store round(100 * cgrid) of the reference's float grid. The reference cannot travel to the GPU
box, so the vectors are committed.  Re-run:  python tests/golden/make_mapgrid_golden.py"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, "..", "shims"))
sys.path.insert(0, "/root/reference")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")

import numpy as np  # noqa: E402

from yag_slam_b200 import karto_compat, synth  # noqa: E402

mod = types.ModuleType("karto_scanmatcher")
mod.ScanMatcherConfig = karto_compat.ScanMatcherConfig
sys.modules["karto_scanmatcher"] = mod
from yag_slam import helpers  # noqa: E402  (reference, unmodified)

out = {}
world = synth.make_world()
img, off = synth.occupancy_image(world, 0.05)
cases = {"world_r05_s05": (img[100:360, 300:560].copy(), 0.05, 0.05), "world_r05_s10": (img[0:200, 0:300].copy(), 0.05, 0.10)}
rng = np.random.default_rng(8)
blobs = np.full((150, 210), 255, np.uint8)
for _ in range(60):
    x, y = rng.integers(0, 205), rng.integers(0, 145)
    blobs[y:y + rng.integers(1, 6), x:x + rng.integers(1, 6)] = 0 if rng.random() < 0.7 else 200
cases["blobs_r02_s06"] = (blobs, 0.02, 0.06)
for name, (m, res, smear) in cases.items():
    cg = helpers.occupancy_grid_map_to_correlation_grid(m, res, smear, 0)
    q = np.round(100.0 * cg)
    assert np.abs(100.0 * cg - q - 0.5).min() > 1e-6  # no value sits on a rounding tie
    out[name + "_img"] = m
    out[name + "_grid"] = q.astype(np.uint8)
    out[name + "_params"] = np.array([res, smear])
    print(name, m.shape, "occupied", int((m == 0).sum()), "nonzero grid cells", int((q > 0).sum()))
np.savez_compressed(os.path.join(HERE, "mapgrid_golden.npz"), **out)
