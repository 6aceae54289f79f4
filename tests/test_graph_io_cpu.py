"""CPU tests of the checkpoint reader / results log: a graph written by the
REFERENCE's own GraphSlam.to_file is read into flat arrays, turned into the re-match batch, and the
file with the results log appended still loads in the unmodified reference. Compute here is the
oracle (this checks the data path, not the kernels)."""
import numpy as np

from oracle import oracle
from yag_slam_b200 import _capi, graph_io, synth

from test_host_cpu import reference_modules  # noqa: F401  (fixture: reference modules on the compat layer)


def _run_reference_slam(reference_modules, world, n=7, P=360, L=4):
    gs, models, sm, serde = reference_modules
    lp = synth.laser_params(P)
    rng = np.random.default_rng(21)
    path = synth.loop_path(n, step=0.2)
    odom = synth.noisy_odometry(path, rng, 0.01, 0.005)
    slam = gs.GraphSlam(sm.Scan2DMatcherCpp({}), None, scan_buffer_len=L)
    live = []
    for k in range(n):
        scan = models.LocalizedRangeScan(synth.cast_scan(world, path[k], P, rng), lp[0], lp[1], lp[2], lp[3], lp[4], lp[5],
                                         *odom[k])
        res, _ = slam.process_scan(scan)
        live.append(res)
    return gs, slam, live


def test_checkpoint_reader_and_results_log_round_trip(reference_modules, world, tmp_path):
    gs, slam, live = _run_reference_slam(reference_modules, world)
    path = tmp_path / "graph.bin"
    slam.to_file(str(path))
    g = graph_io.load(str(path))
    verts = slam.graph.vertices
    assert g.n == len(verts) == 7 and g.scan_buffer_len == 4 and g.loop_matcher_config is None
    assert g.seq_matcher_config["resolution"] == 0.01 and g.loop_search_min_chain_size == 10
    for i, v in enumerate(verts):
        s = v.obj
        assert g.num[i] == s.num and (g.ranges[i] == np.asarray(s.ranges)).all()
        assert g.laser[i].tolist() == [s.min_angle, s.angle_increment, s.min_range, s.range_threshold]
        assert abs(g.corrected[i, 0] - s.corrected_pose.x) == 0 and abs(g.corrected[i, 2] - s.corrected_pose.euler[-1]) < 1e-12
        assert abs(g.odom[i, 1] - s.odom_pose.y) == 0
    assert g.edges.tolist() == [[e.source.obj.num, e.target.obj.num] for e in slam.graph.edges]
    assert g.results is None

    # the re-match batch reproduces what process_scan matched live (same base sets, same initial guess)
    b = graph_io.rematch_batch(g, "odom")
    assert b["base_ptr"].tolist() == [0, 1, 3, 6, 10, 14, 18] and b["base_idx"][-4:].tolist() == [2, 3, 4, 5]
    rec = oracle.match_batch(g.seq_matcher_config, b["pool"], b["starts"], b["counts"], b["query_scan"], b["query_pose"],
                             b["base_ptr"], b["base_idx"], True, True)
    for k in range(1, 7):
        # live poses went through tiny_tf's quaternion algebra, the batch through yaw algebra: equal to rounding
        assert abs(rec[k - 1, 1] - live[k].best_pose.x) < 1e-6 and abs(rec[k - 1, 2] - live[k].best_pose.y) < 1e-6
        assert abs(rec[k - 1, 0] - live[k].response) < 2e-3

    # results log: same framing, loads back bit-exactly, and the reference still reads the file
    records = np.zeros(len(rec), dtype=_capi.RESULT_DTYPE)
    records["response"], records["x"], records["y"], records["heading"] = rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3]
    records["cov"] = rec[:, 4:13]
    blob = graph_io.dumps_with_results(g, records, b, "odom", True, True)
    g2 = graph_io.loads(blob)
    r = g2.results
    assert (r["records"].view(np.uint8) == records.view(np.uint8)).all() and r["guess"] == "odom" and r["do_fine"] is True
    assert (r["base_idx"] == b["base_idx"]).all() and (r["query"] == np.arange(1, 7)).all()
    slam2 = gs.GraphSlam.unbinarize(blob)  # unmodified reference reader ignores the extra key
    assert len(slam2.graph.vertices) == 7 and len(slam2.graph.edges) == len(slam.graph.edges)
    assert slam2.graph.vertices[3].obj.corrected_pose.x == verts[3].obj.corrected_pose.x


def test_stored_guess_and_argument_check(reference_modules, world):
    gs, slam, _ = _run_reference_slam(reference_modules, world, n=4)
    g = graph_io.loads(slam.binarize())
    b = graph_io.rematch_batch(g, "stored")
    assert (b["query_pose"] == g.corrected[1:]).all()
    import pytest
    with pytest.raises(ValueError):
        graph_io.rematch_batch(g, "nope")


def test_committed_checkpoint_reads_without_the_reference():
    import os
    g = graph_io.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_checkpoint.bin"))
    assert g.n == 12 and g.scan_buffer_len == 5 and len(g.ranges[0]) == 360 and g.edges.shape == (11, 2)
    b = graph_io.rematch_batch(g, "odom")
    assert len(b["query_scan"]) == 11 and b["base_ptr"][-1] == len(b["base_idx"]) == 1 + 2 + 3 + 4 + 5 * 7
    rec = oracle.match_batch(g.seq_matcher_config, b["pool"], b["starts"], b["counts"], b["query_scan"], b["query_pose"],
                             b["base_ptr"], b["base_idx"], True, True, 2)
    # re-matching from the odometry guess lands on the corrected poses the mapping run stored
    assert np.abs(rec[:, 1] - g.corrected[1:, 0]).max() < 1e-6 and np.abs(rec[:, 2] - g.corrected[1:, 1]).max() < 1e-6
