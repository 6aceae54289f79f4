"""CPU tests of the host-side mirror of the reference interface: karto-compatible value types,
config plumbing, the Transform shim, and -- when /root/reference is present -- the reference's
own models.py / serde.py / graph_slam.py running UNMODIFIED on top of the compat module
(Wrapper compute swapped for the oracle: this checks the API surface, not the kernels)."""
import os
import sys
import types

import numpy as np
import pytest

from oracle.oracle import KartoOracle
from yag_slam_b200 import karto_compat, scan_matching, synth
from yag_slam_b200.tf import Transform

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def test_config_attribute_set_is_the_reference_key_set():
    cfg = karto_compat.ScanMatcherConfig()
    public = sorted(v for v in dir(cfg) if v[0] != "_")
    assert public == sorted(["angle_variance_penalty", "distance_variance_penalty", "coarse_search_angle_offset",
                             "coarse_angle_resolution", "fine_search_angle_resolution", "use_response_expansion",
                             "range_threshold", "minimum_angle_penalty", "search_size", "resolution",
                             "smear_deviation"])
    assert type(cfg).__name__ == "ScanMatcherConfig"
    c = scan_matching.make_config({"resolution": 0.05, "search_size": 4.0})
    assert c.resolution == 0.05 and c.smear_deviation == 0.05
    with pytest.raises(AssertionError):
        scan_matching.make_config({"smear_deviation": 0.2})
    assert scan_matching.default_config_loop["search_size"] == 4.0


def test_value_types_and_point_reading_cache():
    cfg = karto_compat.LaserScanConfig(-1.0, 1.0, 0.01, 0.05, 30, 20, "")
    s = karto_compat.LocalizedRangeScan(cfg, [1.0, 2.0, 25.0], karto_compat.Pose2(0, 0, 0),
                                        karto_compat.Pose2(1, 2, 0.5), 3, 0.0)
    p1 = s.point_readings()
    assert p1.shape == (2, 2) and s.point_readings() is p1
    s.corrected_pose = karto_compat.Pose2(0, 0, 0)
    p2 = s.point_readings()
    assert p2 is not p1 and p2[0, 0] == 1.0 * np.cos(-1.0)
    s.num = 7
    assert s.num == 7 and s.sensor_pose() == (0.0, 0.0, 0.0)


def test_transform_shim_algebra():
    a = Transform.from_xyt(1.0, 2.0, 0.3)
    b = Transform.from_xyt(-0.5, 0.25, -1.1)
    d = b - a
    c = a + d
    assert abs(c.x - b.x) < 1e-12 and abs(c.y - b.y) < 1e-12 and abs(c.euler[-1] - b.euler[-1]) < 1e-12
    # graph_slam.py:320-322: last.corrected + (query.odom - last.odom) == query.odom when corrected == odom
    assert abs((a + (b - a)).x - b.x) < 1e-12
    p = karto_compat.Pose2(3.0, -1.0, 0.7)
    t = Transform.from_pose2d(p)
    assert (t.x, t.y) == (3.0, -1.0) and abs(t.euler[-1] - 0.7) < 1e-12
    t2 = Transform(**{k: getattr(t, k) for k in ("x", "y", "z", "qx", "qy", "qz", "qw")})
    assert abs(t2.euler[-1] - 0.7) < 1e-12


class _OracleWrapper(object):
    """TEST-ONLY: same interface as karto_compat.Wrapper, compute by the CPU oracle."""

    def __init__(self, config):
        self.config = config
        self._o = KartoOracle(config._as_dict())

    def match_scan(self, query, base_scans, penalty=True, do_fine=False):
        r, p, c = self._o.match(query.point_readings(), query.sensor_pose(),
                                [b.point_readings() for b in base_scans], penalty, do_fine)
        return karto_compat.MatchResult(r, c, karto_compat.Pose2(*p))


@pytest.fixture()
def reference_modules(monkeypatch):
    if not os.path.isdir(os.path.join(REF, "yag_slam")):
        pytest.skip("/root/reference not present (GPU box)")
    mod = types.ModuleType("karto_scanmatcher")
    for n in ("Pose2", "LaserScanConfig", "LocalizedRangeScan", "ScanMatcherConfig", "create_occupancy_grid"):
        setattr(mod, n, getattr(karto_compat, n))
    mod.Wrapper = _OracleWrapper
    monkeypatch.setitem(sys.modules, "karto_scanmatcher", mod)
    monkeypatch.syspath_prepend(os.path.join(HERE, "shims"))
    monkeypatch.syspath_prepend(REF)
    for k in [k for k in sys.modules if k == "yag_slam" or k.startswith("yag_slam.")]:
        monkeypatch.delitem(sys.modules, k)
    import yag_slam.graph_slam as gs
    import yag_slam.models as models
    import yag_slam.scan_matching as sm
    import yag_slam.serde as serde
    yield gs, models, sm, serde
    for k in [k for k in sys.modules if k == "yag_slam" or k.startswith("yag_slam.")]:
        sys.modules.pop(k, None)


def test_reference_consumers_run_unmodified_on_the_compat_module(reference_modules, world):
    gs, models, sm, serde = reference_modules
    P = 360
    lp = synth.laser_params(P)
    rng = np.random.default_rng(3)
    path = synth.loop_path(6, step=0.2)
    odom = synth.noisy_odometry(path, rng, 0.01, 0.005)
    seq = sm.Scan2DMatcherCpp({})
    slam = gs.GraphSlam(seq, None, scan_buffer_len=4)
    for k in range(6):
        r = synth.cast_scan(world, path[k], P, rng)
        scan = models.LocalizedRangeScan(r, lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], *odom[k])
        res, closed = slam.process_scan(scan)
        if k:
            assert 0.0 < res.response <= 1.0 and np.array(res.covariance).shape == (3, 3)
            assert hasattr(res.best_pose, "euler")
    assert len(slam.graph.vertices) == 6 and len(slam.graph.edges) == 5
    # corrected poses stay near the truth
    err = [np.hypot(v.obj.corrected_pose.x - path[i, 0], v.obj.corrected_pose.y - path[i, 1])
           for i, v in enumerate(slam.graph.vertices)]
    assert max(err) < 0.15
    # checkpoint round trip through the reference's serde (class-name keyed)
    blob = slam.binarize()
    slam2 = gs.GraphSlam.unbinarize(blob)
    assert len(slam2.graph.vertices) == 6
    assert slam2.seq_matcher.config.resolution == seq.config.resolution
    d = serde._serialize(seq.config)
    assert d["___name"] == "ScanMatcherConfig" and d["search_size"] == 0.5


def test_numba_twin_agrees_on_the_winning_pose(reference_modules, world):
    """Secondary cross-check: the reference's approximate numba matcher and the
    oracle must pick the same pose to within one coarse cell (2 cm) / one coarse angle step on
    clean synthetic data."""
    gs, models, sm, serde = reference_modules
    P = 360
    lp = synth.laser_params(P)
    rng = np.random.default_rng(9)
    base_pose = synth.loop_path(3)[1]
    true_q = base_pose + np.array([0.06, -0.04, 0.02])
    rb = synth.cast_scan(world, base_pose, P, None)
    rq = synth.cast_scan(world, true_q, P, None)
    base = models.LocalizedRangeScan(rb, lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], *base_pose)
    query = models.LocalizedRangeScan(rq, lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], *base_pose)
    py = sm.Scan2DMatcherPy({})
    r_py = py.match_scan(query, [base], penalty=False, do_fine=False)
    cpp = sm.Scan2DMatcherCpp({})
    r_or = cpp.match_scan(query, [base], False, False)
    assert abs(r_py.best_pose.x - r_or.best_pose.x) <= 0.0201
    assert abs(r_py.best_pose.y - r_or.best_pose.y) <= 0.0201
    assert abs(r_py.best_pose.euler[-1] - r_or.best_pose.euler[-1]) <= 0.0350
    assert abs(r_or.best_pose.x - true_q[0]) < 0.08 and abs(r_or.best_pose.y - true_q[1]) < 0.08


def test_matcher_golden_pins_the_oracle():
    g = np.load(os.path.join(HERE, "golden", "matcher_golden.npz"))
    from oracle import oracle
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_matcher_golden as mk
    for name, (cfg, kw, pen, fine) in mk.CASES.items():
        out = oracle.match_batch(cfg, g[f"{name}_pool"], g[f"{name}_starts"], g[f"{name}_counts"],
                                 g[f"{name}_query_scan"], g[f"{name}_query_pose"], g[f"{name}_base_ptr"],
                                 g[f"{name}_base_idx"], pen, fine, 2)
        assert (out.view(np.uint64) == g[f"{name}_ref"].view(np.uint64)).all(), name
    o = KartoOracle()
    r, p, c = o.match(g["testpy_query"], g["testpy_pose"], [g["testpy_base"]], True, True)
    assert (np.concatenate([[r], p, c.ravel()]) == g["testpy_ref"]).all()
