"""BASELINE cfg 2 as SURVEY.md 8(d) defines it: the reference's UNMODIFIED GraphSlam.process_scan
(yag_slam/graph_slam.py:306-339) driving a `karto_scanmatcher` stand-in over a synthetic 720-beam trajectory.

The reference package is taken from baseline/_ref (pip-installed there from /root/reference, see DESIGN.md;
it travels to the GPU box) or, in the build container, straight from /root/reference. The absent `tiny_tf` and
`sba_cpp` wheels are replaced by the test-only stand-ins in tests/shims (graph optimisation is out of scope:
SPA2d.compute is a no-op)."""
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_path():
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "yag_slam")):
            return p
    return None


def import_reference(wrapper_cls=None):
    """Fresh import of the reference's graph_slam / models / scan_matching with `karto_scanmatcher` bound to
    yag_slam_b200.karto_compat (wrapper_cls: replaces its Wrapper, e.g. by an oracle-backed one in tests)."""
    ref = reference_path()
    if ref is None:
        raise RuntimeError("reference package not found (baseline/_ref or /root/reference)")
    from yag_slam_b200 import karto_compat
    mod = types.ModuleType("karto_scanmatcher")
    for n in ("Pose2", "LaserScanConfig", "LocalizedRangeScan", "ScanMatcherConfig", "create_occupancy_grid", "Wrapper"):
        setattr(mod, n, getattr(karto_compat, n))
    if wrapper_cls is not None:
        mod.Wrapper = wrapper_cls
    sys.modules["karto_scanmatcher"] = mod
    for p in (os.path.join(ROOT, "tests", "shims"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    for k in [k for k in sys.modules if k == "yag_slam" or k.startswith("yag_slam.")]:
        del sys.modules[k]
    import yag_slam.graph_slam as gs
    import yag_slam.models as models
    import yag_slam.scan_matching as sm
    return gs, models, sm


def make_trajectory(world, n, beams, seed=2, step=0.25):
    """SURVEY 8d cfg 2: n poses on the closed loop, odometry = truth + cumulative N(0, 0.02 m / 0.01 rad),
    ranges cast with N(0, 0.01 m) noise."""
    from yag_slam_b200 import synth
    rng = np.random.default_rng(seed)
    path = synth.loop_path(n, step=step)
    odom = synth.noisy_odometry(path, rng, 0.02, 0.01)
    ranges = [synth.cast_scan(world, path[k], beams, rng) for k in range(n)]
    return path, odom, ranges


def run_sequential(mods, world, n, beams, seed=2, with_loop=False, scan_buffer_len=10, traj=None):
    """Feeds the trajectory to GraphSlam.process_scan. Returns dict(poses [n][3], response [n], closed,
    match_s [n-1] = wall clock of each seq_matcher.match_scan, total_s = wall clock of the whole loop)."""
    from yag_slam_b200 import synth
    gs, models, sm = mods
    path, odom, ranges = traj if traj is not None else make_trajectory(world, n, beams, seed)
    lp = synth.laser_params(beams)
    seq = sm.Scan2DMatcherCpp({})
    loop = sm.Scan2DMatcherCpp({}, loop=True) if with_loop else None
    slam = gs.GraphSlam(seq, loop, scan_buffer_len=scan_buffer_len)
    match_s = []
    inner = seq.match_scan

    def timed_match(*a, **k):
        t0 = time.perf_counter()
        r = inner(*a, **k)
        match_s.append(time.perf_counter() - t0)
        return r

    seq.match_scan = timed_match
    poses = np.zeros((n, 3))
    resp = np.zeros(n)
    closed = 0
    t0 = time.perf_counter()
    for k in range(n):
        scan = models.LocalizedRangeScan(ranges[k], lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], *odom[k])
        res, cl = slam.process_scan(scan)
        if res is not None:
            resp[k] = res.response
        closed += 1 if cl else 0
        cp = scan.corrected_pose
        poses[k] = (cp.x, cp.y, cp.euler[-1])
    total = time.perf_counter() - t0
    return dict(poses=poses, response=resp, closed=closed, match_s=np.array(match_s), total_s=total, truth=path,
                n_vertices=len(slam.graph.vertices), n_edges=len(slam.graph.edges))
