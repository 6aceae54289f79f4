"""BASELINE cfg 2 as defined: the reference's own, UNMODIFIED GraphSlam.process_scan
(yag_slam/graph_slam.py:306-339, imported from baseline/_ref) drives karto_compat.Wrapper -- i.e. the CUDA path
through the C ABI -- and must reproduce, pose for pose and bit for bit, the same loop driven by the CPU oracle."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _OracleWrapper(object):
    """TEST-ONLY: karto_compat.Wrapper's interface, compute by the CPU oracle."""

    def __init__(self, config):
        from oracle.oracle import KartoOracle
        self.config = config
        self._o = KartoOracle(config._as_dict())

    def match_scan(self, query, base_scans, penalty=True, do_fine=False):
        from yag_slam_b200 import karto_compat
        r, p, c = self._o.match(query.point_readings(), query.sensor_pose(),
                                [b.point_readings() for b in base_scans], penalty, do_fine)
        return karto_compat.MatchResult(r, c, karto_compat.Pose2(*p))


@pytest.mark.parametrize("with_loop", [False, True])
def test_reference_graphslam_on_the_cuda_wrapper_matches_the_oracle_run(with_loop):
    from harness import refslam
    from yag_slam_b200 import synth
    if refslam.reference_path() is None:
        pytest.skip("reference package not installed (baseline/_ref)")
    world = synth.make_world()
    # with the loop matcher: a full lap plus a little (the loop is 71 m long; chains need >= 10 consecutive scans
    # within sqrt(3) m of the query, graph_slam.py:291) so that loop closures happen
    n, beams, step = (262, 360, 0.3) if with_loop else (60, 720, 0.25)
    traj = refslam.make_trajectory(world, n, beams, seed=2, step=step)
    gpu = refslam.run_sequential(refslam.import_reference(), world, n, beams, with_loop=with_loop, traj=traj)
    cpu = refslam.run_sequential(refslam.import_reference(_OracleWrapper), world, n, beams, with_loop=with_loop, traj=traj)
    assert gpu["n_vertices"] == cpu["n_vertices"] == n and gpu["n_edges"] == cpu["n_edges"]
    assert gpu["closed"] == cpu["closed"]
    assert (gpu["response"].view(np.uint64) == cpu["response"].view(np.uint64)).all()
    assert (gpu["poses"].view(np.uint64) == cpu["poses"].view(np.uint64)).all()
    if with_loop:
        assert gpu["closed"] >= 1, "the trajectory should close its loop at least once"
    err = np.hypot(gpu["poses"][:, 0] - gpu["truth"][:, 0], gpu["poses"][:, 1] - gpu["truth"][:, 1])
    assert np.median(err) < 0.15
