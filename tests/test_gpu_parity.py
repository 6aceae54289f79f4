"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs. Bars (BASELINE.json north_star):
  * response and best pose: BIT-EXACT (the response sums are integers; every transcendental
    is evaluated by the same libm on both sides);
  * covariance: relative error <= 1e-5 (COV_RTOL below; the positional sums are reduced in
    parallel on the GPU, sequentially on the CPU);
  * intermediate products (smear kernel, correlation grid bytes, lookup-offset tables): exact.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
COV_RTOL = 1e-5
LOOP = dict(search_size=4.0, resolution=0.05)


def _matcher(cfg=None, **kw):
    from yag_slam_b200.matcher import ScanMatcherB200
    return ScanMatcherB200(cfg, **kw)


def _run(m, b, penalty, do_fine):
    return m.match_pool(b["pool"], b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"],
                        b["base_idx"], penalty, do_fine)


def _assert_parity(out, ref, what=""):
    ref = np.asarray(ref).reshape(-1, 13)
    assert len(out) == len(ref)
    assert (out["status"] == 0).all()
    for k, col in (("response", 0), ("x", 1), ("y", 2), ("heading", 3)):
        bad = np.where(out[k].view(np.uint64) != ref[:, col].copy().view(np.uint64))[0]
        assert len(bad) == 0, f"{what}: {k} differs at matches {bad[:8]}: gpu {out[k][bad[:4]]} ref {ref[bad[:4], col]}"
    cov, rc = out["cov"], ref[:, 4:]
    err = np.abs(cov - rc)
    # element-wise relative bound; the XY term is a cancelling sum (exactly 0 for a symmetric
    # response surface), so it is bounded relative to sqrt(XX * YY) instead of to itself
    scale = np.abs(rc).copy()
    scale[:, 1] = scale[:, 3] = np.maximum(scale[:, 1], np.sqrt(np.abs(rc[:, 0] * rc[:, 4])))
    tol = COV_RTOL * scale
    assert (err <= tol).all(), f"{what}: covariance rel err {np.max(err / np.maximum(scale, 1e-300)):.3e}"


def test_smear_kernel_exact():
    from oracle.oracle import KartoOracle
    for cfg in (None, LOOP, dict(search_size=0.3, smear_deviation=0.07), dict(smear_deviation=0.03)):
        m = _matcher(cfg, max_slots=1)
        o = KartoOracle(cfg)
        d, od = m.dims(), o.dims()
        for k in ("side", "margin", "roi", "half_kernel", "kernel_size", "border", "width", "height", "stride"):
            assert d[k] == od[k], k
        assert d["grid_bytes"] == od["data_size"]
        assert (m.debug_kernel() == o.kernel()).all()
        m.close()


def test_correlation_grid_bytes_and_offset_tables_exact(world):
    import scenarios
    from oracle.oracle import KartoOracle
    from yag_slam_b200 import _capi
    for cfg, P, nb, seed in ((None, 360, 1, 21), (None, 720, 10, 22), (LOOP, 720, 10, 23),
                             (dict(search_size=0.3, smear_deviation=0.07), 500, 3, 24)):
        b = scenarios.make_batch(world, 3, P, nb, seed)
        m = _matcher(cfg, max_slots=4)
        m.set_debug(_capi.DEBUG_KEEP_GRIDS)
        _run(m, b, False, False)
        o = KartoOracle(cfg)
        for i in range(3):
            bases = [b["points"][s] for s in b["base_idx"][b["base_ptr"][i]:b["base_ptr"][i + 1]]]
            ref_grid = o.build_grid(b["query_pose"][i], bases)
            got = m.debug_grid(i)
            assert ref_grid.max() == 100
            assert (got == ref_grid).all(), f"grid bytes differ: {np.argwhere(got != ref_grid)[:5]}"
            q = b["points"][b["query_scan"][i]]
            pose = b["query_pose"][i]
            cfgd = o.params
            ref_off = o.compute_offsets(q, pose, pose[2], cfgd.coarse_search_angle_offset, cfgd.coarse_angle_resolution)
            got_off = m.debug_offsets(i)
            assert got_off.shape == ref_off.shape and (got_off == ref_off).all()
        # after the debug flag is dropped the slots are cleared again: a different batch is exact
        m.set_debug(0)
        b2 = scenarios.make_batch(world, 3, P, nb, seed + 100)
        _assert_parity(_run(m, b2, True, True), scenarios.oracle_results(cfg, b2, True, True), "after-clear")
        m.close()


def test_committed_golden_vectors():
    g = np.load(os.path.join(HERE, "golden", "matcher_golden.npz"))
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_matcher_golden as mk
    for name, (cfg, kw, pen, fine) in mk.CASES.items():
        m = _matcher(cfg, max_slots=8)
        b = {k: g[f"{name}_{k}"] for k in ("pool", "starts", "counts", "query_scan", "query_pose", "base_ptr", "base_idx")}
        _assert_parity(_run(m, b, pen, fine), g[f"{name}_ref"], name)
        m.close()
    # the reference's own smoke geometry (test.py:23-43)
    from yag_slam_b200.matcher import pack_pool
    pool, starts, counts = pack_pool([g["testpy_query"], g["testpy_base"]])
    m = _matcher(None, max_slots=1)
    out = m.match_pool(pool, starts, counts, [0], g["testpy_pose"][None, :], [0, 1], [1], True, True)
    _assert_parity(out, g["testpy_ref"][None, :], "test.py geometry")
    m.close()


@pytest.mark.parametrize("penalty,do_fine", [(True, True), (True, False), (False, True), (False, False)])
def test_seq_config_vs_oracle(world, penalty, do_fine):
    import scenarios
    b = scenarios.make_batch(world, 16, 720, 10, 31)
    m = _matcher(None, max_slots=16)
    _assert_parity(_run(m, b, penalty, do_fine), scenarios.oracle_results(None, b, penalty, do_fine), "seq")
    m.close()


def test_loop_config_with_degenerate_chains_and_expansion(world):
    import scenarios
    b = scenarios.make_batch(world, 40, 720, 10, 4, perturb=(1.0, 0.2), degenerate_frac=0.25)
    m = _matcher(LOOP)
    out = _run(m, b, False, False)
    _assert_parity(out, scenarios.oracle_results(LOOP, b, False, False), "loop")
    assert (out["n_passes"] == 4).sum() >= 4  # degenerate chains ran all three expansions
    assert (out["n_ties"][out["n_passes"] == 4] == 41 * 41 * 81).all()
    # expansion disabled: a single pass for the same chains
    cfg = dict(LOOP, use_response_expansion=False)
    m2 = _matcher(cfg)
    out2 = _run(m2, b, False, True)
    _assert_parity(out2, scenarios.oracle_results(cfg, b, False, True), "loop-noexp")
    m.close()
    m2.close()


def test_zero_row_pruning_is_exact(world):
    """The throughput sweep skips (point, lattice row) spans whose tiles hold no non-zero cell;
    results must equal the unpruned sweep and the oracle for the sequential config (26 x 26
    lattice), the loop config (41 x 41: two row groups, two column chunks), penalised or not,
    including query points whose windows leave the grid (no pruning, flat bounds check)."""
    import scenarios
    from yag_slam_b200 import _capi
    cases = ((None, 720, 10, 31, (0.1, 0.05), True, 20.0), (LOOP, 720, 10, 32, (1.0, 0.2), False, 20.0),
             (dict(search_size=0.3, smear_deviation=0.07), 360, 2, 33, (0.1, 0.05), True, 20.0),
             # the query scans keep readings up to 30 m while the matcher grid only spans 12 m
             (dict(range_threshold=12.0), 360, 3, 34, (0.1, 0.05), True, 30.0))
    for cfg, P, nb, seed, perturb, penalty, rthr in cases:
        b = scenarios.make_batch(world, 48, P, nb, seed, perturb=perturb, degenerate_frac=0.1, range_threshold=rthr)
        ref = scenarios.oracle_results(cfg, b, penalty, False)
        m = _matcher(cfg, max_slots=48)
        a = _run(m, b, penalty, False).copy()
        m.set_debug(_capi.DEBUG_NO_PRUNE)
        c = _run(m, b, penalty, False).copy()
        m.close()
        _assert_parity(a, ref, "pruned sweep %r" % (cfg,))
        _assert_parity(c, ref, "unpruned sweep %r" % (cfg,))
        assert a.tobytes() == c.tobytes()


def test_latency_path_device_chained_fine_pass(world):
    """Small batches take the latency path: one host->device copy, the fine pass chained on the
    device behind the coarse pass, one synchronisation. Results must equal the general path
    (DEBUG_NO_SPECULATE) and the oracle, including matches that cannot be speculated (tied coarse
    winners along a wall, empty grids that trigger response expansion) and fall back per match."""
    import scenarios
    from yag_slam_b200 import _capi
    from yag_slam_b200.matcher import pack_pool
    for cfg, n, P, nb, seed, degen in ((None, 1, 360, 1, 41, 0.0), (None, 5, 720, 10, 42, 0.0),
                                       (None, 8, 360, 3, 43, 0.4), (dict(search_size=0.3, smear_deviation=0.07), 3, 500, 2, 44, 0.0)):
        b = scenarios.make_batch(world, n, P, nb, seed, perturb=(0.07, 0.03), degenerate_frac=degen)
        ref = scenarios.oracle_results(cfg, b, True, True)
        m = _matcher(cfg, max_slots=8)
        a = _run(m, b, True, True).copy()
        w = m.last_work()
        spec = w["speculative_fine_passes"]
        # one query: the resident kernel; a handful: the single cooperative kernel
        assert (w["resident_requests"] == 1) if n == 1 else (w["latency_kernel_launches"] == 1), \
            "the latency path did not run"
        m.set_debug(_capi.DEBUG_NO_MEGA)
        a2 = _run(m, b, True, True).copy()
        w2 = m.last_work()
        assert w2["latency_kernel_launches"] == 0 and w2["speculative_fine_passes"] == spec
        m.set_debug(_capi.DEBUG_NO_MEGA | _capi.DEBUG_NO_SPECULATE)
        c = _run(m, b, True, True).copy()
        assert m.last_work()["speculative_fine_passes"] == 0
        m.set_debug(_capi.DEBUG_NO_SPECULATE)
        c2 = _run(m, b, True, True).copy()
        coarse_only = _run(m, b, True, False).copy()
        m.close()
        _assert_parity(a, ref, "latency kernel")
        _assert_parity(c, ref, "general path")
        assert a2.tobytes() == c.tobytes() == c2.tobytes()
        if n > 1:
            assert a.tobytes() == a2.tobytes()
        else:
            # the resident kernel reduces the A.9 covariance sums over a different tree: same pose / response
            # bits, covariance equal to rounding
            for k in ("response", "x", "y", "heading", "n_passes", "n_ties"):
                assert (a[k] == a2[k]).all(), k
            assert np.allclose(a["cov"], a2["cov"], rtol=1e-12, atol=0)
        _assert_parity(coarse_only, scenarios.oracle_results(cfg, b, True, False), "latency kernel, coarse only")
        if degen == 0.0:
            assert spec >= 1, "the device-chained fine pass never ran"
    # a featureless wall: every pose along it ties, so the coarse pass has many winners
    from oracle.oracle import KartoOracle
    xs = np.linspace(3.0, -3.0, 301)  # counter-clockwise as seen from the origin (FindValidPoints keeps it)
    wall = np.stack([xs, np.full_like(xs, 2.0)], axis=1)
    pool, starts, counts = pack_pool([wall, wall])
    args = (pool, starts, counts, np.array([0], np.int32), np.array([[0.0, 0.0, 0.0]]), np.array([0, 1], np.int32),
            np.array([1], np.int32))
    m = _matcher(None, max_slots=2)
    out = m.match_pool(*args, True, True)
    assert m.last_work()["speculative_fine_passes"] == 0 and out["n_passes"][0] == 2
    r, pose, cov = KartoOracle(None).match(wall, (0.0, 0.0, 0.0), [wall], True, True)
    _assert_parity(out, np.concatenate([[r], pose, cov.reshape(-1)]), "tied wall")
    m.close()


def test_heading_wraparound_on_every_path(world):
    """Queries heading along -x (heading = -pi on the top edge of the loop, guesses on both sides of
    the +-pi cut): the coarse / fine search angles leave [-pi, pi], so NormalizeAngle changes them
    (separate cos/sin of the normalised headings for the tie average, atan2 of the averaged heading
    lands on either side of the cut). Latency kernel, general path and throughput path vs the oracle."""
    import scenarios
    from yag_slam_b200 import _capi
    for n, slots in ((1, 4), (6, 8), (200, 0)):
        b = scenarios.make_batch(world, n, 360, 3, 71 + n, perturb=(0.07, 0.05), path_start=38.0,
                                 path_step=0.25 if n <= 6 else 0.07)  # (every pose stays on the top edge)
        assert (np.abs(np.abs(b["query_pose"][:, 2]) - np.pi) < 0.06).all()
        if n > 1:
            assert (b["query_pose"][:, 2] < -np.pi).any() and (b["query_pose"][:, 2] > -np.pi).any()
        ref = scenarios.oracle_results(None, b, True, True)
        m = _matcher(None, max_slots=slots)
        a = _run(m, b, True, True).copy()
        w = m.last_work()
        lat = w["latency_kernel_launches"]
        assert (w["resident_requests"] == 1) if n == 1 else (lat == (1 if n <= 6 else 0))
        _assert_parity(a, ref, "wrap n=%d" % n)
        if n <= 6:
            m.set_debug(_capi.DEBUG_NO_MEGA | _capi.DEBUG_NO_SPECULATE)
            c = _run(m, b, True, True).copy()
            assert a.tobytes() == c.tobytes()
        m.close()
        assert (np.abs(a["heading"]) > 3.0).all()  # matched headings stay next to the cut, on either side


def test_long_base_chains(world):
    """Base sets much longer than the running-scan buffer (loop-closure chains have no upper length,
    graph_slam.py:274-304): 70 scans per match exceeds the latency kernel's 64-scan limit (general
    path), 40 x 8 matches sits inside it, and 70 x 40 matches runs the throughput path with more cells
    per match than the 16-bit per-tile candidate counters take (searched build)."""
    import scenarios
    for n, nb, slots, lat in ((2, 70, 4, 0), (8, 40, 8, 1), (40, 70, 0, 0)):
        b = scenarios.make_batch(world, n, 360, nb, 300 + n, perturb=(0.07, 0.03), path_step=0.1)
        ref = scenarios.oracle_results(None, b, True, True)
        m = _matcher(None, max_slots=slots)
        out = _run(m, b, True, True).copy()
        assert m.last_work()["latency_kernel_launches"] == lat
        _assert_parity(out, ref, "long chains n=%d nb=%d" % (n, nb))
        assert _run(m, b, True, True).tobytes() == out.tobytes()  # slots were cleared
        m.close()


def test_lanes_split_large_batches(world):
    """Large batches are split over internal lanes (own slots / stream / host thread); results are
    those of the single-lane path and of the oracle, with a host pool and with a device pool."""
    import torch
    import scenarios
    b = scenarios.make_batch(world, 300, 360, 3, 61, perturb=(0.1, 0.05), degenerate_frac=0.05)
    ref = scenarios.oracle_results(None, b, True, True)
    m1 = _matcher(None, max_slots=128, lanes=1)
    a = _run(m1, b, True, True).copy()
    assert m1.last_work()["lanes"] == 0
    m1.close()
    m3 = _matcher(None, max_slots=192, lanes=3)
    c = _run(m3, b, True, True).copy()
    assert m3.last_work()["lanes"] == 3 and m3.dims()["slots"] == 192
    dpool = torch.from_numpy(b["pool"]).cuda()
    d = m3.match_pool(dpool, b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"],
                      True, True).copy()
    m3.close()
    _assert_parity(a, ref, "one lane")
    _assert_parity(c, ref, "three lanes")
    assert a.tobytes() == c.tobytes() == d.tobytes()


def test_shared_query_and_multiwave(world):
    import scenarios
    b = scenarios.make_batch(world, 32, 720, 10, 5, perturb=(1.0, 0.2), shared_query=True)
    ref = scenarios.oracle_results(LOOP, b, False, False)
    m = _matcher(LOOP)
    _assert_parity(_run(m, b, False, False), ref, "shared")
    m.close()
    m = _matcher(LOOP, max_slots=5)  # 7 waves
    _assert_parity(_run(m, b, False, False), ref, "shared-multiwave")
    m.close()


def test_ragged_and_empty_inputs(world):
    import scenarios
    from yag_slam_b200.matcher import pack_pool
    b = scenarios.make_batch(world, 3, 360, 2, 41)
    pts = list(b["points"])
    empty = len(pts)
    pts.append(np.zeros((0, 2)))
    short = len(pts)
    pts.append(pts[0][:7].copy())
    pool, starts, counts = pack_pool(pts)
    q = b["query_scan"]
    query_scan = [q[0], empty, q[1], short, q[2]]
    poses = np.array([b["query_pose"][0], [1.0, 2.0, 0.3], b["query_pose"][1], b["query_pose"][0], b["query_pose"][2]])
    base_ptr = [0, 2, 4, 4, 6, 9]  # third match has NO base scans; last has 3 incl. an empty one
    bi = b["base_idx"]
    base_idx = [bi[0], bi[1], bi[0], bi[1], bi[2], bi[3], bi[4], empty, bi[5]]
    bb = dict(pool=pool, starts=starts, counts=counts, query_scan=np.array(query_scan, np.int32), query_pose=poses,
              base_ptr=np.array(base_ptr, np.int32), base_idx=np.array(base_idx, np.int32))
    m = _matcher(None, max_slots=2)
    out = _run(m, bb, True, True)
    _assert_parity(out, scenarios.oracle_results(None, bb, True, True), "ragged")
    assert out["response"][1] == 0.0 and out["n_passes"][1] == 0 and out["cov"][1][0] == 500.0  # empty query
    assert out["n_passes"][2] == 5  # no base points: coarse + 3 expansions + fine
    assert len(m.match_pool(pool, starts, counts, [], np.zeros((0, 3)), [0], [], True, True)) == 0
    m.close()


def test_ros_node_config_and_other_resolutions(world):
    import scenarios
    for cfg, P in ((dict(search_size=0.3, smear_deviation=0.07), 720), (dict(resolution=0.02, search_size=0.6), 360),
                   (dict(smear_deviation=0.03, coarse_angle_resolution=0.02, fine_search_angle_resolution=0.002), 500)):
        b = scenarios.make_batch(world, 6, P, 5, 51)
        m = _matcher(cfg, max_slots=6)
        _assert_parity(_run(m, b, True, True), scenarios.oracle_results(cfg, b, True, True), str(cfg))
        m.close()


def test_parameter_errors():
    with pytest.raises(ValueError):
        _matcher(dict(smear_deviation=0.2))  # > 10 * resolution (Karto runtime_error)
    with pytest.raises(ValueError):
        _matcher(dict(smear_deviation=0.001))


def test_dropin_api_single_and_batch(world):
    """Scan2DMatcherCpp.match_scan / match_scan_batch (reference scan_matching.py:32-42) with
    yag_slam.models.LocalizedRangeScan-like objects (`._scan` is the karto-compatible scan)."""
    from oracle.oracle import KartoOracle
    from yag_slam_b200 import karto_compat as kc
    from yag_slam_b200 import synth
    from yag_slam_b200.scan_matching import Scan2DMatcherCpp

    class Scan(object):  # the part of yag_slam/models.py the matcher touches
        def __init__(self, ranges, lp, pose):
            cfg = kc.LaserScanConfig(lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], "")
            self._scan = kc.LocalizedRangeScan(cfg, ranges, kc.Pose2(*pose), kc.Pose2(*pose), 0, 0.0)

    P = 360
    lp = synth.laser_params(P)
    rng = np.random.default_rng(61)
    path = synth.loop_path(12)
    scans = [Scan(synth.cast_scan(world, p, P, rng), lp, p + rng.normal(0, [0.03, 0.03, 0.01])) for p in path]
    seq = Scan2DMatcherCpp({}, max_slots=8)
    o = KartoOracle()
    singles = []
    for k in range(4, 12):
        r = seq.match_scan(scans[k], scans[k - 4:k], True, True)
        er, ep, ec = o.match(scans[k]._scan.point_readings(), scans[k]._scan.sensor_pose(),
                             [s._scan.point_readings() for s in scans[k - 4:k]], True, True)
        assert r.response == er and r.meta is None
        assert (r.best_pose.x, r.best_pose.y) == (ep[0], ep[1]) and abs(r.best_pose.euler[-1] - ep[2]) < 1e-12
        assert np.allclose(np.array(r.covariance), ec, rtol=COV_RTOL, atol=0)
        assert np.linalg.inv(np.array(r.covariance)).shape == (3, 3) and r.covariance[0][0] > 0
        singles.append(r)
    batch = seq.match_scan_batch(scans[4:12], [scans[k - 4:k] for k in range(4, 12)], True, True)
    for a, b in zip(singles, batch):
        assert a.response == b.response and a.best_pose.x == b.best_pose.x and a.best_pose.y == b.best_pose.y
    loop = Scan2DMatcherCpp({}, loop=True, max_slots=8)
    assert loop.config.resolution == 0.05 and loop.config.search_size == 4.0
    r = loop.match_scan(scans[11], scans[0:10], False, False)
    assert 0 < r.response <= 1
    with pytest.raises(TypeError):
        Scan2DMatcherCpp(None)  # reference: dict.update(None) raises (scan_matching.py:36)


def test_sequential_mapping_loop_matches_oracle(world):
    """cfg-2 shaped driver: loop-carried running-scan matching (graph_slam.py:306-339 without the
    graph bookkeeping), GPU vs oracle pose-for-pose."""
    from oracle.oracle import KartoOracle
    from yag_slam_b200 import _capi, synth
    from yag_slam_b200.matcher import pack_pool
    P, n = 720, 40
    lp = synth.laser_params(P)
    rng = np.random.default_rng(2)
    path = synth.loop_path(n)
    odom = synth.noisy_odometry(path, rng)
    ranges = [synth.cast_scan(world, p, P, rng) for p in path]
    m = _matcher(None, max_slots=1)
    o = KartoOracle()

    def run(match_fn):
        corrected = [odom[0].copy()]
        pts = [_capi.point_readings(ranges[0], lp[0], lp[2], lp[3], lp[5], *corrected[0])]
        resp = []
        for k in range(1, n):
            d = odom[k] - odom[k - 1]
            c, s = np.cos(odom[k - 1][2]), np.sin(odom[k - 1][2])
            loc = np.array([c * d[0] + s * d[1], -s * d[0] + c * d[1], d[2]])
            c, s = np.cos(corrected[-1][2]), np.sin(corrected[-1][2])
            guess = corrected[-1] + np.array([c * loc[0] - s * loc[1], s * loc[0] + c * loc[1], loc[2]])
            q = _capi.point_readings(ranges[k], lp[0], lp[2], lp[3], lp[5], *guess)
            r, pose = match_fn(q, guess, pts[-10:])
            resp.append(r)
            corrected.append(np.array(pose))
            pts.append(_capi.point_readings(ranges[k], lp[0], lp[2], lp[3], lp[5], *pose))
        return np.array(resp), np.array(corrected)

    def gpu_fn(q, guess, bases):
        pool, starts, counts = pack_pool([q] + list(bases))
        r = m.match_pool(pool, starts, counts, [0], guess[None, :], [0, len(bases)], np.arange(1, len(bases) + 1), True, True)[0]
        return r["response"], (r["x"], r["y"], r["heading"])

    def cpu_fn(q, guess, bases):
        r, p, _ = o.match(q, guess, bases, True, True)
        return r, p

    gr, gc = run(gpu_fn)
    cr, cc = run(cpu_fn)
    assert (gr == cr).all() and (gc == cc).all()
    assert np.abs(gc[:, :2] - path[:, :2]).max() < 0.25
    m.close()


def test_fullsize_loop_closure_batch_properties(world):
    """BASELINE cfg 3 size: one query vs 4,096 candidate chains x 10 scans (P=720, loop config,
    expansion on, 10% degenerate). Size-independent properties + an oracle-checked sample."""
    import scenarios
    n = 4096
    b = scenarios.make_batch(world, n, 720, 10, 3, perturb=(1.0, 0.2), degenerate_frac=0.1, shared_query=True)
    m = _matcher(LOOP)
    out = _run(m, b, False, False)
    assert m.launch_count() > 0 and (out["status"] == 0).all()
    # (a) permutation invariance: shuffling the chains permutes the results, bit for bit
    perm = np.random.default_rng(0).permutation(n)
    nb = np.diff(b["base_ptr"])
    base_ptr = np.concatenate([[0], np.cumsum(nb[perm])]).astype(np.int32)
    base_idx = np.concatenate([b["base_idx"][b["base_ptr"][i]:b["base_ptr"][i + 1]] for i in perm]).astype(np.int32)
    bp = dict(b, query_scan=b["query_scan"][perm], query_pose=b["query_pose"][perm], base_ptr=base_ptr, base_idx=base_idx)
    outp = _run(m, bp, False, False)
    assert outp.tobytes() == out[perm].tobytes()
    # (b) idempotence / determinism: a second run is byte-identical (grids were cleared correctly)
    assert _run(m, b, False, False).tobytes() == out.tobytes()
    # (c) chains that are the same scan set give the same record (the generator cycles 7 sets)
    key = [tuple(b["base_idx"][b["base_ptr"][i]:b["base_ptr"][i + 1]]) for i in range(n)]
    first = {}
    for i, k in enumerate(key):
        if k in first:
            assert out[i].tobytes() == out[first[k]].tobytes()
        else:
            first[k] = i
    # (d) base-scan order does not matter (the smear is a pure max)
    rev = np.concatenate([b["base_idx"][b["base_ptr"][i]:b["base_ptr"][i + 1]][::-1] for i in range(64)]).astype(np.int32)
    br = dict(b, query_scan=b["query_scan"][:64], query_pose=b["query_pose"][:64], base_ptr=b["base_ptr"][:65], base_idx=rev)
    assert _run(m, br, False, False).tobytes() == out[:64].tobytes()
    # (e) every distinct chain set against the oracle
    idx = np.array(sorted(first.values()))[:16]
    sub_ptr = np.concatenate([[0], np.cumsum(nb[idx])]).astype(np.int32)
    sub_idx = np.concatenate([b["base_idx"][b["base_ptr"][i]:b["base_ptr"][i + 1]] for i in idx]).astype(np.int32)
    bs = dict(b, query_scan=b["query_scan"][idx], query_pose=b["query_pose"][idx], base_ptr=sub_ptr, base_idx=sub_idx)
    _assert_parity(out[idx], scenarios.oracle_results(LOOP, bs, False, False), "cfg3 sample")
    # responses are multiples of 1 / (P * 100) (integer sums), no penalty
    P = b["counts"][b["query_scan"][0]]
    q = out["response"] * (P * 100)
    assert np.abs(q - np.round(q)).max() < 1e-6
    m.close()


def test_fullsize_sequential_log_rematch_properties(world):
    """BASELINE cfg 2 size: 2,000 scans x 720 beams re-matched against their 10 running scans
    (the bench workload). Wave-size independence + an oracle-checked sample."""
    import scenarios
    n = 2000
    b = scenarios.make_batch(world, n, 720, 10, 2, perturb=(0.1, 0.05))
    m = _matcher(None)
    out = _run(m, b, True, True)
    m.close()
    m2 = _matcher(None, max_slots=333)
    out2 = _run(m2, b, True, True)
    m2.close()
    assert out.tobytes() == out2.tobytes()
    idx = np.arange(0, n, 67)
    nb = np.diff(b["base_ptr"])
    sub_ptr = np.concatenate([[0], np.cumsum(nb[idx])]).astype(np.int32)
    sub_idx = np.concatenate([b["base_idx"][b["base_ptr"][i]:b["base_ptr"][i + 1]] for i in idx]).astype(np.int32)
    bs = dict(b, query_scan=b["query_scan"][idx], query_pose=b["query_pose"][idx], base_ptr=sub_ptr, base_idx=sub_idx)
    _assert_parity(out[idx], scenarios.oracle_results(None, bs, True, True), "cfg2 sample")
    assert (out["response"] > 0.5).mean() > 0.95  # the log re-matches well


def test_highres_config_one_match(world):
    """BASELINE cfg 4 (P=1081, res 0.005, search 1.0, fine 0.00175): with smear 0.045 (K=37) and
    with yag_slam's default smear 0.05 = 10 * res (K=41, 68 MB grid), which is Karto's
    order-dependent regime (the stamp is 100 on its four edge neighbours too, so AddScan skips
    points whose cell an earlier point already saturated)."""
    import scenarios
    b = scenarios.make_batch(world, 2, 1081, 1, 4, perturb=(0.1, 0.03))
    for smear, K in ((0.045, 37), (0.05, 41)):
        cfg = dict(resolution=0.005, search_size=1.0, fine_search_angle_resolution=0.00175, smear_deviation=smear)
        m = _matcher(cfg, max_slots=2)
        assert m.dims()["roi"] == 8201 and m.dims()["kernel_size"] == K
        _assert_parity(_run(m, b, True, True), scenarios.oracle_results(cfg, b, True, True), "cfg4 smear %g" % smear)
        m.close()


def test_wide_smear_ordered_stamps(world):
    """smear_deviation = 10 * resolution at the default resolution, 10 base scans per match: many
    points fall on cells an earlier point already saturated, so the grid depends on Karto's
    processing order. Grid bytes and match results must equal the sequential oracle."""
    import scenarios
    from oracle.oracle import KartoOracle
    from yag_slam_b200 import _capi
    cfg = dict(smear_deviation=0.1)
    b = scenarios.make_batch(world, 6, 720, 10, 52, perturb=(0.1, 0.05))
    m = _matcher(cfg, max_slots=8)
    m.set_debug(_capi.DEBUG_KEEP_GRIDS)
    out = _run(m, b, True, True).copy()
    o = KartoOracle(cfg)
    assert (o.kernel() == 100).sum() == 5
    for i in range(6):
        bases = [b["points"][s] for s in b["base_idx"][b["base_ptr"][i]:b["base_ptr"][i + 1]]]
        ref_grid = o.build_grid(b["query_pose"][i], bases)
        got = m.debug_grid(i)
        assert (got == ref_grid).all(), f"grid bytes differ: {np.argwhere(got != ref_grid)[:5]}"
    m.set_debug(0)
    _assert_parity(out, scenarios.oracle_results(cfg, b, True, True), "wide smear")
    big = scenarios.make_batch(world, 400, 720, 10, 53, perturb=(0.1, 0.05))
    _assert_parity(_run(m, big, True, False), scenarios.oracle_results(cfg, big, True, False), "wide smear batch")
    m.close()


def test_device_resident_pool_and_stream(world):
    import torch
    import scenarios
    b = scenarios.make_batch(world, 12, 720, 10, 71)
    m = _matcher(None, max_slots=12)
    ref = _run(m, b, True, True)
    dpool = torch.from_numpy(b["pool"]).cuda()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        out = m.match_pool(dpool, b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"],
                           b["base_idx"], True, True, stream=st.cuda_stream)
    assert out.tobytes() == ref.tobytes()
    pinned = torch.from_numpy(b["pool"]).pin_memory()
    out = m.match_pool(pinned, b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"],
                       b["base_idx"], True, True)
    assert out.tobytes() == ref.tobytes()
    m.close()


def test_single_rank_nccl_gather(world):
    import torch
    import torch.distributed as dist
    import scenarios
    from yag_slam_b200 import distributed
    b = scenarios.make_batch(world, 9, 360, 3, 81)
    m = _matcher(None, max_slots=9)
    ref = _run(m, b, True, True)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda:0"))
    try:
        full = distributed.match_pool_sharded(m.match_pool, b["pool"], b["starts"], b["counts"], b["query_scan"],
                                              b["query_pose"], b["base_ptr"], b["base_idx"], True, True,
                                              device=torch.device("cuda:0"))
    finally:
        dist.destroy_process_group()
    assert full.tobytes() == ref.tobytes()
    m.close()


def test_wave_sliced_pool_upload(world):
    """Lanes + host-resident pool: the pool is uploaded wave by wave (merged contiguous scan ranges, an
    event behind every wave). Log-shaped access (scan k vs scans k-3..k-1: few ranges per wave), a
    shuffled pool (fragmented: falls back to one copy) and a pinned pool all give the single-lane records."""
    import torch
    from oracle import oracle
    from yag_slam_b200 import synth
    from yag_slam_b200.matcher import pack_pool
    n, P, L = 401, 180, 3
    rng = np.random.default_rng(77)
    path = synth.loop_path(n, step=0.2)
    guess = path + np.concatenate([rng.uniform(-0.08, 0.08, (n, 2)), rng.uniform(-0.03, 0.03, (n, 1))], axis=1)
    base_pts = [scenarios_scan_points(world, path[i], P, rng) for i in range(n)]
    query_pts = [scenarios_scan_points(world, path[i], P, rng, guess[i]) for i in range(n)]
    order = np.arange(2 * n)
    for shuffled in (False, True):
        if shuffled:
            order = rng.permutation(2 * n)
        inv = np.argsort(order)  # scan id -> position in the pool
        allpts = base_pts + query_pts
        pool, starts_p, counts_p = pack_pool([allpts[j] for j in order])
        starts, counts = starts_p[inv], counts_p[inv]  # indexed by scan id
        qs = np.arange(n + 1, 2 * n, dtype=np.int32)
        bp, bi = [0], []
        for k in range(1, n):
            bi.extend(range(max(0, k - L), k))
            bp.append(len(bi))
        args = (starts, counts, qs, guess[1:], np.array(bp, np.int32), np.array(bi, np.int32), True, True)
        m1 = _matcher(None, max_slots=64, lanes=1)
        a = m1.match_pool(pool, *args).copy()
        m1.close()
        m2 = _matcher(None, max_slots=128, lanes=2)
        c = m2.match_pool(pool, *args).copy()
        assert m2.last_work()["lanes"] == 2
        if not shuffled:  # only what the matches reference crosses the bus (scan 0's query copy never does)
            assert m2.last_work()["h2d_bytes"] < pool.nbytes + 4 * 1024 * 1024
        pinned = torch.from_numpy(pool).pin_memory()
        d = m2.match_pool(pinned, *args).copy()
        m2.close()
        assert a.tobytes() == c.tobytes() == d.tobytes(), "shuffled=%s" % shuffled
    ref = oracle.match_batch(None, pool, starts, counts, qs, guess[1:], np.array(bp, np.int32), np.array(bi, np.int32),
                             True, True)
    _assert_parity(a, ref, "sliced upload")


def scenarios_scan_points(world, pose, P, rng, sense_at=None):
    # points of a scan taken at `pose`, expressed at `sense_at` (the query's initial guess) when given
    from yag_slam_b200 import synth
    if sense_at is None:
        return synth.scan_points(world, pose, P, rng)
    return synth.scan_points(world, sense_at, P, rng, sense_pose=pose)


def test_exact_candidate_lists_equal_the_searched_build(world):
    """The grid build with exact per-tile candidate lists (k_find_valid's 16-bit counters / cursors), stamped by
    k_tile_stamp_lists (per-column-group / per-tile-half step lists) or by k_tile_stamp's per-candidate loop
    (YSM_DEBUG_NO_HALF_LISTS), gives byte-identical grids and identical records to the build that searches the
    cell list per tile (YSM_DEBUG_NO_CANDLISTS), on the sequential and the loop configuration."""
    import scenarios
    from yag_slam_b200 import _capi
    # (the last configuration has a 33 x 33 smear kernel: the widest class k_tile_stamp_lists takes)
    for cfg, nb, P in ((None, 10, 720), (LOOP, 10, 360), (dict(resolution=0.02, smear_deviation=0.05), 4, 360),
                       (dict(smear_deviation=0.08), 3, 360)):
        b = scenarios.make_batch(world, 40, P, nb, 91, perturb=(0.1, 0.05), degenerate_frac=0.1)
        grids, recs = [], []
        for flags in (0, _capi.DEBUG_NO_HALF_LISTS, _capi.DEBUG_NO_CANDLISTS):
            m = _matcher(cfg, max_slots=48, lanes=1)
            m.set_debug(flags | _capi.DEBUG_KEEP_GRIDS)
            recs.append(_run(m, b, True, True).copy())
            grids.append([m.debug_grid(i) for i in (0, 7, 39)])
            m.close()
        assert recs[0].tobytes() == recs[1].tobytes() == recs[2].tobytes()
        for ga, gb, gc in zip(*grids):
            assert (ga == gb).all() and (ga == gc).all()
        assert max(int(ga.max()) for ga in grids[0]) == 100  # (a degenerate match has an empty grid)
        _assert_parity(recs[0], scenarios.oracle_results(cfg, b, True, True), "candidate lists")


def test_beams_without_an_in_range_reading(world):
    """Karto's MatchScan returns early only for a scan WITHOUT range readings. A query that has beams but none
    inside [min_range, range_threshold] runs the whole schedule on an empty lookup table: every pose of every
    pass ties at response 0 (all three response expansions, then the fine pass). ysm_batch::scan_raw_count carries
    the raw beam count; the result must equal the oracle's run of that schedule, bit for bit."""
    import scenarios
    from oracle.oracle import KartoOracle
    from yag_slam_b200 import karto_compat as kc
    from yag_slam_b200 import synth
    from yag_slam_b200.matcher import pack_pool
    rng = np.random.default_rng(9)
    base = synth.scan_points(world, (0.0, 0.0, 0.0), 360, rng)
    pool, starts, counts = pack_pool([np.zeros((0, 2)), base])
    one = (np.array([0], np.int32), np.array([[1.0, -2.0, 0.3]]), np.array([0, 1], np.int32), np.array([1], np.int32))
    for cfg in (None, LOOP, dict(use_response_expansion=False)):
        m = _matcher(cfg, max_slots=2)
        o = KartoOracle(cfg)
        for fine in (False, True):
            out = m.match_pool(pool, starts, counts, *one, True, fine, scan_raw_count=np.array([360, 360], np.int32))
            r, p, cov = o.match(np.zeros((0, 2)), (1.0, -2.0, 0.3), [base], True, fine, n_raw=360)
            _assert_parity(out, np.concatenate([[r], p, cov.reshape(-1)]), "beams without readings %r fine=%s" % (cfg, fine))
            assert out["n_passes"][0] == (4 if (cfg or {}).get("use_response_expansion", True) else 1) + int(fine)
            # no beams at all (or no raw counts given): MatchScan's early return
            for raw in (np.array([0, 360], np.int32), None):
                e = m.match_pool(pool, starts, counts, *one, True, fine, scan_raw_count=raw)
                assert e["response"][0] == 0.0 and (e["x"][0], e["y"][0], e["heading"][0]) == (1.0, -2.0, 0.3)
                assert e["cov"][0][0] == 500.0 and e["n_passes"][0] == 0
        m.close()
    # through the reference-facing wrapper: every beam beyond the range threshold
    lp = synth.laser_params(360)
    cfg = kc.LaserScanConfig(lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], "")
    b = kc.LocalizedRangeScan(cfg, synth.cast_scan(world, (0.0, 0.0, 0.0), 360, rng), kc.Pose2(0, 0, 0), kc.Pose2(0, 0, 0), 0, 0.0)
    q = kc.LocalizedRangeScan(cfg, np.full(360, 25.0), kc.Pose2(1.0, -2.0, 0.3), kc.Pose2(1.0, -2.0, 0.3), 1, 0.0)
    w = kc.Wrapper(kc.ScanMatcherConfig())
    res = w.match_scan(q, [b], True, True)
    r, p, cov = KartoOracle(None).match(np.zeros((0, 2)), (1.0, -2.0, 0.3), [b.point_readings()], True, True, n_raw=360)
    assert res.response == r and (res.best_pose.x, res.best_pose.y, res.best_pose.yaw) == tuple(p)
    assert np.allclose(res.covariance, cov, rtol=1e-12)
