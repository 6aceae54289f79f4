"""GPU parity tests of create_occupancy_grid (csrc/ysm_occ.cu) against the CPU oracle
(oracle/occgrid_oracle.c): dimensions / offset / pass and hit counters / image bit-exact."""
import math

import numpy as np
import pytest

from yag_slam_b200 import occupancy, raytracing, synth

pytestmark = pytest.mark.gpu


def _both(log, res, thr):
    from oracle import oracle
    g = occupancy.occupancy_grid_from_arrays(log["poses"], log["lasers"], log["ranges"], log["beam_ptr"], res, thr)
    o = oracle.occupancy_grid(log["poses"], log["lasers"], log["ranges"], log["beam_ptr"], res, thr)
    return g, o


def _assert_same(g, o):
    assert (g.width, g.height) == (o["width"], o["height"])
    assert g.offset.x == o["offset_x"] and g.offset.y == o["offset_y"]  # bit-exact (libm re-evaluation)
    p, h = g.counts()
    assert (p == o["passes"]).all(), "pass counters differ in %d cells" % int((p != o["passes"]).sum())
    assert (h == o["hits"]).all()
    assert (g.image == o["image"]).all()


@pytest.mark.parametrize("n_scans,n_beams,res,thr", [(1, 360, 0.05, 12.0), (40, 720, 0.05, 12.0),
                                                      (25, 1081, 0.02, 20.0), (60, 360, 0.1, 6.0)])
def test_occupancy_grid_vs_oracle(world, n_scans, n_beams, res, thr):
    log = synth.make_scan_log(world, n_scans, n_beams, seed=11 + n_scans)
    g, o = _both(log, res, thr)
    _assert_same(g, o)
    assert g.info["launches"] >= 4 and g.info["rays"] == n_scans * n_beams
    vals = set(np.unique(g.image).tolist())
    assert vals <= {0, 200, 255} and (n_scans < 3 or {0, 255} <= vals)


def test_defective_readings_and_hand_case(world):
    log = synth.make_scan_log(world, 30, 720, seed=5, defects=0.08)
    g, o = _both(log, 0.05, 12.0)
    _assert_same(g, o)
    # hand case of tests/test_occupancy_cpu.py
    laser = np.array([[-math.pi, 2 * math.pi / 4, 0.05, 30.0]])
    g = occupancy.occupancy_grid_from_arrays([(0.0, 0.0, 0.0)], laser, [3.0, 2.0, 5.0, 4.0], [0, 4], 1.0, 10.0)
    assert (g.width, g.height) == (8, 6)
    p, h = g.counts()
    assert p[2, 3] == 4 and h.sum() == 2 and p[2, 0] == 2


def test_exactness_guard_on_axis_aligned_walls():
    # noise-free scans of an axis-aligned room from grid-aligned poses: many end points sit exactly
    # on cell rounding boundaries and on the faces of the bounding box -- the cases the host libm
    # re-evaluation exists for
    w = synth.make_world(n_pillars=0)
    n, nb = 12, 720
    poses = np.array([(-5.0 + 0.5 * i, 0.25 * i - 1.0, 0.0) for i in range(n)])
    lp = synth.laser_params(nb)
    ranges = np.concatenate([synth.cast_scan(w, p, nb, None) for p in poses])
    log = dict(poses=poses, lasers=np.tile([lp[0], lp[2], lp[3], lp[4]], (n, 1)), ranges=ranges,
               beam_ptr=(np.arange(n + 1) * nb).astype(np.int32))
    g, o = _both(log, 0.05, 25.0)
    _assert_same(g, o)
    assert g.info["box_candidates"] > 100  # every wall point lies on a face of the box


def test_dropin_create_occupancy_grid_and_raywalk_chain(world):
    from yag_slam_b200 import karto_compat as kc
    from oracle import oracle
    log = synth.make_scan_log(world, 50, 360, seed=3)
    lp = synth.laser_params(360)
    cfg = kc.LaserScanConfig(lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], "laser")
    scans = []
    for i in range(50):
        p = kc.Pose2(*log["poses"][i])
        scans.append(kc.LocalizedRangeScan(cfg, log["ranges"][i * 360:(i + 1) * 360].tolist(), p, p, i, 0.0))
    grid = kc.create_occupancy_grid(scans, 0.05, 12)
    o = oracle.occupancy_grid(log["poses"], log["lasers"], log["ranges"], log["beam_ptr"], 0.05, 12.0)
    assert grid.image.shape == (grid.height, grid.width) and (grid.image == o["image"]).all()
    assert grid.offset.x == o["offset_x"] and grid.offset.y == o["offset_y"]
    assert kc.create_occupancy_grid([], 0.05, 12) is None
    # ray-walk straight from the HBM-resident image == ray-walk of the host copy == oracle
    sx, sy = (log["poses"][10, 0] - grid.offset.x) / 0.05, (log["poses"][10, 1] - grid.offset.y) / 0.05
    ang = np.arange(-180, 180, 1.0)
    a = raytracing.raytrace_many(grid, ang, [(sx, sy)])
    b = raytracing.raytrace_many(grid.image, ang, [(sx, sy)])
    c = oracle.raywalk_sweep(grid.image, ang, sx, sy)
    assert (a.view(np.uint32) == b.view(np.uint32)).all() and (a[0].view(np.uint32) == c.view(np.uint32)).all()


def test_fullsize_map_properties(world):
    # BASELINE cfg-2/5 shape: the 2,000-scan 720-beam log at 0.05 m/px, threshold 12 m. Size-independent
    # properties: sum(pass) = in-bounds Bresenham cells + valid end points, sum(hit) = valid in-bounds
    # end points; building from two halves and adding the counters equals the whole (linearity).
    log = synth.make_scan_log(world, 2000, 720, seed=2)
    g = occupancy.occupancy_grid_from_arrays(log["poses"], log["lasers"], log["ranges"], log["beam_ptr"], 0.05, 12.0)
    p, h = g.counts()
    r = log["ranges"]
    valid = (r > 0.05) & (r < 30.0) & (r < 12.0 - 1e-6)
    assert h.sum() <= valid.sum() and h.sum() >= 0.98 * valid.sum()
    assert p.sum() > h.sum() and p.sum() <= g.info["cells_visited"] + h.sum()
    assert (h <= p).all()
    assert 0 < g.info["cell_fixups"] < 1000  # ~2e-6 of the rays sit within 1e-6 cell of a rounding boundary
    from oracle import oracle
    k = 150  # oracle on a prefix at full beam count
    sub = dict(poses=log["poses"][:k], lasers=log["lasers"][:k], ranges=r[:k * 720], beam_ptr=log["beam_ptr"][:k + 1])
    gs, os_ = _both(sub, 0.05, 12.0)
    _assert_same(gs, os_)
