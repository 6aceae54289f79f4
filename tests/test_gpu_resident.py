"""GPU parity tests of the resident single-query kernel (k_match_resident, csrc/ysm_resident.cuh):
ysm_match_batch with ONE match -- the reference's own call pattern, Wrapper.match_scan once per scan
(yag_slam/scan_matching.py:40-42, graph_slam.py:220,236,326) -- against the CPU oracle. Bars as in
test_gpu_parity.py: response / pose bit-exact, covariance <= 1e-5 relative."""
import os
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
LOOP = dict(search_size=4.0, resolution=0.05)


@pytest.fixture(scope="module")
def world():
    from yag_slam_b200 import synth
    return synth.make_world()


def _single(m, b, i, penalty, do_fine):
    from yag_slam_b200.distributed import slice_batch
    qs, qp, bp, bi = slice_batch(b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"], i, i + 1)
    return m.match_pool(b["pool"], b["starts"], b["counts"], qs, qp, bp, bi, penalty, do_fine)


def _check(m, cfg, b, penalty, do_fine, what, expect_resident=True, sleep=0.0):
    import scenarios
    from test_gpu_parity import _assert_parity
    ref = scenarios.oracle_results(cfg, b, penalty, do_fine)
    served = 0
    for i in range(len(ref)):
        out = _single(m, b, i, penalty, do_fine)
        served += m.last_work()["resident_requests"]
        _assert_parity(out, ref[i:i + 1], "%s, match %d" % (what, i))
        if sleep:
            time.sleep(sleep)
    if expect_resident:
        assert served == len(ref), "%s: only %d of %d queries were served by the resident kernel" % (what, served, len(ref))
    return served


def test_single_queries_bit_exact(world):
    """cfg 1 / cfg-2 shapes, with and without penalty / fine pass, back to back on one handle (every request
    reuses the same grid slot, tile ownership and shared-memory state of the resident kernel)."""
    import scenarios
    from yag_slam_b200.matcher import ScanMatcherB200
    m = ScanMatcherB200(None, max_slots=4, lanes=1)
    _check(m, None, scenarios.make_batch(world, 24, 360, 1, 101, perturb=(0.07, 0.03)), True, True, "cfg1")
    _check(m, None, scenarios.make_batch(world, 16, 720, 10, 102), True, True, "cfg2 shape")
    _check(m, None, scenarios.make_batch(world, 8, 720, 10, 103), False, False, "coarse only, no penalty")
    _check(m, None, scenarios.make_batch(world, 8, 500, 3, 104), False, True, "no penalty, fine")
    _check(m, None, scenarios.make_batch(world, 6, 230, 1, 105, perturb=(0.05, 0.02)), True, False, "penalty, coarse only")
    assert m.launch_count() <= 8, "the kernel did not stay resident between back-to-back requests"
    m.close()


def test_other_configurations(world):
    import scenarios
    from yag_slam_b200.matcher import ScanMatcherB200
    for cfg, P, nb, seed, pert in ((LOOP, 720, 10, 111, (1.0, 0.2)), (dict(search_size=0.3, smear_deviation=0.07), 500, 2, 112, (0.05, 0.03)),
                                   (dict(smear_deviation=0.03), 360, 4, 113, (0.07, 0.03))):
        m = ScanMatcherB200(cfg, max_slots=2, lanes=1)
        # (coarse-resolution grids can overflow a CTA's tile lists: those requests take the general path)
        strict = cfg is not LOOP
        s1 = _check(m, cfg, scenarios.make_batch(world, 8, P, nb, seed, perturb=pert), False, False, "cfg %r coarse" % (cfg,),
                    expect_resident=strict)
        s2 = _check(m, cfg, scenarios.make_batch(world, 6, P, nb, seed + 50, perturb=pert), True, True, "cfg %r fine" % (cfg,),
                    expect_resident=strict)
        assert s1 + s2 >= 7
        m.close()


def test_idle_exit_and_relaunch(world):
    """The kernel leaves the device after resident_idle_us without a request; the next call relaunches it."""
    import scenarios
    from yag_slam_b200.matcher import ScanMatcherB200
    m = ScanMatcherB200(None, max_slots=2, lanes=1, resident_idle_us=300)
    b = scenarios.make_batch(world, 6, 360, 2, 121)
    _check(m, None, b, True, True, "idle exit", sleep=0.005)
    assert m.launch_count() >= 5, "the kernel never left the device"
    m.close()
    m = ScanMatcherB200(None, max_slots=2, lanes=1, resident_idle_us=-1)  # resident path disabled
    assert _check(m, None, b, True, True, "resident disabled", expect_resident=False) == 0
    m.close()


def test_fallback_cases(world):
    """Requests the kernel cannot finish exactly are rerun through the general path: empty grids (response
    expansion), tied coarse winners along a featureless wall, an empty query."""
    import scenarios
    from oracle.oracle import KartoOracle
    from test_gpu_parity import _assert_parity
    from yag_slam_b200.matcher import ScanMatcherB200, pack_pool
    m = ScanMatcherB200(LOOP, max_slots=2, lanes=1)
    b = scenarios.make_batch(world, 12, 720, 10, 131, perturb=(1.0, 0.2), degenerate_frac=0.4)
    _check(m, LOOP, b, False, False, "degenerate chains", expect_resident=False)
    m.close()
    m = ScanMatcherB200(None, max_slots=2, lanes=1)
    xs = np.linspace(3.0, -3.0, 301)
    wall = np.stack([xs, np.full_like(xs, 2.0)], axis=1)
    pool, starts, counts = pack_pool([wall, wall, np.zeros((0, 2))])
    one = (np.array([[0.0, 0.0, 0.0]]), np.array([0, 1], np.int32), np.array([1], np.int32))
    out = m.match_pool(pool, starts, counts, np.array([0], np.int32), *one, True, True)
    r, pose, cov = KartoOracle(None).match(wall, (0.0, 0.0, 0.0), [wall], True, True)
    _assert_parity(out, np.concatenate([[r], pose, cov.reshape(-1)]), "tied wall")
    out = m.match_pool(pool, starts, counts, np.array([2], np.int32), *one, True, True)  # empty query
    assert out["response"][0] == 0.0 and out["cov"][0][0] == 500.0
    # and the handle keeps serving ordinary requests afterwards
    _check(m, None, scenarios.make_batch(world, 4, 360, 1, 132, perturb=(0.07, 0.03)), True, True, "after fallbacks")
    m.close()


def test_handles_interleaved_and_batches_between(world):
    """seq matcher, loop matcher, seq matcher ... (the reference's loop-closure flow, graph_slam.py:220,236):
    one resident kernel per device, the handles take turns; a throughput batch in between ends it first."""
    import scenarios
    from test_gpu_parity import _assert_parity, _run
    from yag_slam_b200.matcher import ScanMatcherB200
    seq = ScanMatcherB200(None, max_slots=16, lanes=1)
    loop = ScanMatcherB200(LOOP, max_slots=16, lanes=1)
    b1 = scenarios.make_batch(world, 6, 360, 2, 141)
    b2 = scenarios.make_batch(world, 6, 720, 5, 142, perturb=(1.0, 0.2))
    r1 = scenarios.oracle_results(None, b1, True, True)
    r2 = scenarios.oracle_results(LOOP, b2, False, False)
    for i in range(6):
        _assert_parity(_single(seq, b1, i, True, True), r1[i:i + 1], "seq %d" % i)
        _assert_parity(_single(loop, b2, i, False, False), r2[i:i + 1], "loop %d" % i)
    bb = scenarios.make_batch(world, 12, 360, 3, 143)
    _assert_parity(_run(seq, bb, True, True), scenarios.oracle_results(None, bb, True, True), "batch between singles")
    for i in range(6):
        _assert_parity(_single(seq, b1, i, True, True), r1[i:i + 1], "seq again %d" % i)
    seq.close()
    loop.close()


def test_device_resident_scan_store(world):
    """ysm_batch::scan_tag: the sequential-mapping pattern (scan k against scans k-10 .. k-1, graph_slam.py:326)
    uploads every scan once; later matches read it from the device-resident store. Results stay bit-exact, also
    when the store wraps (more scans than slots), a scan appears twice in one request, or tags are absent."""
    import scenarios
    from test_gpu_parity import _assert_parity
    from yag_slam_b200 import synth
    from yag_slam_b200.matcher import ScanMatcherB200, pack_pool
    rng = np.random.default_rng(151)
    path = synth.loop_path(60, step=0.25)
    pts = [synth.scan_points(world, p, 360, rng) for p in path]
    pool, starts, counts = pack_pool(pts)
    tags = (np.arange(len(pts)) + 1000).astype(np.uint64)
    m = ScanMatcherB200(None, max_slots=2, lanes=1)
    from oracle.oracle import KartoOracle
    o = KartoOracle(None)
    hits = served = 0
    for k in range(10, 60):
        base = np.arange(k - 10, k, dtype=np.int32)
        if k == 30:
            base[3] = base[2]  # the same scan twice in one request
        pose = path[k] + np.array([0.05, -0.03, 0.02])
        out = m.match_pool(pool, starts, counts, np.array([k], np.int32), pose[None, :], np.array([0, 10], np.int32), base,
                           True, True, scan_tag=tags if k % 7 else None)
        w = m.last_work()
        served += w["resident_requests"]  # (a tied coarse winner is finished by the general path)
        hits += w["scan_store_hits"]
        r, p, cov = o.match(pts[k], tuple(pose), [pts[j] for j in base], True, True)
        _assert_parity(out, np.concatenate([[r], p, cov.reshape(-1)]), "scan store, scan %d" % k)
    assert served >= 40, "only %d of 50 queries were finished by the resident kernel" % served
    assert hits >= 250, "the running scans were not served from the device-resident store (%d hits)" % hits
    m.close()


def test_wrapper_match_scan_uses_it(world):
    """The reference-facing call (karto_compat.Wrapper.match_scan) is served by the resident kernel and the
    doorbell round trip is measurable."""
    from oracle.oracle import KartoOracle
    from yag_slam_b200 import karto_compat as kc
    from yag_slam_b200 import synth
    P, nb = 360, 1
    lp = synth.laser_params(P)
    rng = np.random.default_rng(1)
    path = synth.loop_path(nb + 1)
    cfg = kc.LaserScanConfig(lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], "")
    base = [kc.LocalizedRangeScan(cfg, synth.cast_scan(world, p, P, rng), kc.Pose2(*p), kc.Pose2(*p), i, 0.0)
            for i, p in enumerate(path[:nb])]
    q = kc.LocalizedRangeScan(cfg, synth.cast_scan(world, path[0] + np.array([0.07, -0.04, 0.03]), P, rng),
                              kc.Pose2(*path[0]), kc.Pose2(*path[0]), nb, 0.0)
    w = kc.Wrapper(kc.ScanMatcherConfig())
    assert w.matcher.dims()["slots"] <= 8, "a single-query Wrapper must not take the throughput HBM budget"
    ref = KartoOracle(None).match(q.point_readings(), q.sensor_pose(), [s.point_readings() for s in base], True, True)
    for _ in range(5):
        r = w.match_scan(q, base, True, True)
        assert r.response == ref[0] and (r.best_pose.x, r.best_pose.y, r.best_pose.yaw) == tuple(ref[1])
        assert w.matcher.last_work()["resident_requests"] == 1
    assert w.matcher.last_work()["scan_store_hits"] == 2, "query and base scan should come from the scan store"
    # a new corrected pose makes new point readings: new content tag, no stale hit
    q.corrected_pose = kc.Pose2(path[0][0] + 0.01, path[0][1], path[0][2])
    ref2 = KartoOracle(None).match(q.point_readings(), q.sensor_pose(), [s.point_readings() for s in base], True, True)
    r = w.match_scan(q, base, True, True)
    assert r.response == ref2[0] and (r.best_pose.x, r.best_pose.y, r.best_pose.yaw) == tuple(ref2[1])
    assert w.matcher.last_work()["scan_store_hits"] == 1
    rtt = w.matcher.ping(50)
    assert rtt.shape == (50,) and (rtt > 0).all()


def test_dense_tiles_and_stamp_list_overflow(world):
    """The tile collect of the resident kernel bumps a tile's counter once per (warp, tile): base scans whose
    points repeat (every running scan sees the same wall from almost the same pose) put hundreds of stamps into a
    few tiles. Up to 256 stamps per tile the kernel serves the request; past that the CTA reports the overflow at
    barrier 2 and the general path reruns the match. Either way the result is the oracle's."""
    import scenarios
    from oracle.oracle import KartoOracle
    from test_gpu_parity import _assert_parity
    from yag_slam_b200 import synth
    from yag_slam_b200.matcher import ScanMatcherB200, pack_pool
    rng = np.random.default_rng(141)
    pose = np.array(synth.loop_path(40)[7], dtype=np.float64)
    qpose = pose + np.array([0.05, -0.03, 0.02])  # the query: taken at `pose`, localised 6 cm off
    q = synth.scan_points(world, qpose, 720, rng, sense_pose=pose)
    m = ScanMatcherB200(None, max_slots=2, lanes=1)
    o = KartoOracle(None)
    served = []
    for nbase, jitter in ((3, 0.003), (6, 0.002), (12, 0.0)):
        # nbase scans taken within millimetres of each other: their cells coincide or neighbour
        base = [synth.scan_points(world, pose + np.array([jitter * k, -jitter * k, 0.0]), 720, rng) for k in range(nbase)]
        pool, starts, counts = pack_pool([q] + base)
        out = m.match_pool(pool, starts, counts, np.array([0], np.int32), qpose[None, :],
                           np.array([0, nbase], np.int32), np.arange(1, nbase + 1, dtype=np.int32), True, True)
        served.append(m.last_work()["resident_requests"])
        r, p, cov = o.match(q, tuple(qpose), base, True, True)
        _assert_parity(out, np.concatenate([[r], p, cov.reshape(-1)]), "dense tiles, %d coincident base scans" % nbase)
    assert served[0] == 1, "three coincident scans fit the per-tile lists: the resident kernel must serve them"
    # and the handle keeps serving ordinary requests after an overflow
    _check(m, None, scenarios.make_batch(world, 3, 360, 1, 142, perturb=(0.07, 0.03)), True, True, "after dense tiles")
    m.close()
