"""CPU tests of the oracle (oracle/): hand-computable cases for every Karto function it
restates, the derived sizes of Appendix E, and the ray-walk against the
golden vectors generated from the reference's own numba code."""
import math
import os

import numpy as np
import pytest

from oracle import oracle
from oracle.oracle import KartoOracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_sizes_appendix_e():
    d = KartoOracle().dims()
    assert (d["side"], d["margin"], d["roi"], d["kernel_size"]) == (51, 2000, 4051, 21)
    assert (d["width"], d["stride"], d["data_size"]) == (4073, 4080, 16617840)
    d = KartoOracle(dict(search_size=4.0, resolution=0.05)).dims()
    assert (d["side"], d["margin"], d["roi"], d["kernel_size"]) == (81, 400, 881, 5)
    assert (d["width"], d["stride"], d["data_size"]) == (887, 888, 787656)
    d = KartoOracle(dict(search_size=1.0, resolution=0.005, fine_search_angle_resolution=0.00175)).dims()
    assert (d["side"], d["roi"], d["kernel_size"], d["stride"], d["data_size"]) == (201, 8201, 41, 8248, 67988264)
    d = KartoOracle(dict(search_size=0.3, smear_deviation=0.07)).dims()  # ROS-node seq config
    assert (d["side"], d["roi"], d["kernel_size"], d["width"], d["stride"]) == (31, 4031, 29, 4061, 4064)


def test_kernel_values():
    k = KartoOracle().kernel()
    assert k.shape == (21, 21) and k[10, 10] == 100
    assert (k == k.T).all() and (k == k[::-1, ::-1]).all()
    for i, j in [(0, 0), (3, 4), (10, 0), (10, 10), (7, -2)]:
        z = math.exp(-0.5 * (math.hypot(i * 0.01, j * 0.01) / 0.05) ** 2)
        assert k[10 + j, 10 + i] == math.floor(z * 100 + 0.5)
    assert k[10, 0] == 14 and k[0, 0] == 2


def test_smear_bounds_rejected():
    with pytest.raises(RuntimeError):
        KartoOracle(dict(smear_deviation=0.004))  # < 0.5 * res
    with pytest.raises(RuntimeError):
        KartoOracle(dict(smear_deviation=0.11))  # > 10 * res


def test_point_readings_filter_and_values():
    r = np.array([0.01, 1.0, 2.0, 25.0, 20.0])
    p = oracle.point_readings(r, -1.0, 0.5, 0.05, 20.0, 1.0, 2.0, 0.25)
    assert len(p) == 3  # 0.01 < min_range and 25 > threshold dropped, 20.0 kept (inclusive)
    for k, i in enumerate([1, 2, 4]):
        a = 0.25 + -1.0 + i * 0.5
        assert p[k, 0] == 1.0 + r[i] * math.cos(a) and p[k, 1] == 2.0 + r[i] * math.sin(a)


def test_find_valid_points_hand_cases():
    # counter-clockwise wall seen from the origin: points 0.06 m apart, trigger every 2nd point
    xs = np.arange(0, 12) * 0.06
    pts = np.column_stack([np.full_like(xs, 2.0), xs - 0.3])
    m = oracle.find_valid_points(pts, 0.0, 0.0)
    # triggers at 2,4,...,10 ; last segment [10, 12) is never emitted
    assert m.tolist() == [1] * 10 + [0, 0]
    # same wall traversed clockwise -> wrong side, nothing kept
    m = oracle.find_valid_points(pts[::-1].copy(), 0.0, 0.0)
    assert m.sum() == 0
    # all points closer than 10 cm to the first: no trigger at all
    m = oracle.find_valid_points(np.column_stack([np.full(5, 1.0), np.arange(5) * 0.01]), 0.0, 0.0)
    assert m.sum() == 0


def _wall_scan(n=200, dist=3.0, span=2.0):
    ys = np.linspace(-span, span, n)
    return np.column_stack([np.full(n, dist), ys])


def test_identity_match_recovers_pose():
    o = KartoOracle()
    # L-shaped corner gives a unique pose
    a = _wall_scan()
    b = np.column_stack([np.linspace(3.0, -1.0, 200), np.full(200, 2.0)])
    pts = np.vstack([a, b])
    resp, pose, cov = o.match(pts, (0.0, 0.0, 0.0), [pts], True, True)
    assert resp > 0.9
    assert abs(pose[0]) < 1e-12 and abs(pose[1]) < 1e-12 and abs(pose[2]) < 1e-12
    assert cov[0, 0] > 0 and cov[1, 1] > 0 and cov[2, 2] > 0


def test_tie_averaging_along_a_wall():
    # a single straight wall: every y-translation ties; Karto averages them (penalty off)
    o = KartoOracle()
    pts = _wall_scan(400, 3.0, 3.0)
    resp, pose, cov = o.match(pts, (0.0, 0.0, 0.0), [pts], False, False)
    n = o.dims()["last_ties"]
    assert n > 1
    # the average of n lattice positions (step 0.02) is a multiple of 0.02 / n
    q = pose[1] / (0.02 / n)
    assert abs(q - round(q)) < 1e-6 and abs(pose[1]) <= 0.25


def test_empty_query_and_no_overlap():
    o = KartoOracle()
    resp, pose, cov = o.match(np.zeros((0, 2)), (1.0, 2.0, 0.5), [_wall_scan()], True, True)
    assert resp == 0.0 and pose == (1.0, 2.0, 0.5)
    assert cov[0, 0] == 500.0 and cov[1, 1] == 500.0 and cov[2, 2] == 4 * 0.0349 ** 2
    # no base points: best == 0 -> all three response expansions run, every pose ties
    resp, pose, cov = o.match(_wall_scan(), (0.0, 0.0, 0.0), [np.zeros((0, 2))], True, False)
    d = o.dims()
    assert resp == 0.0 and d["last_passes"] == 4
    assert d["n_angles"] == 81 and d["last_ties"] == 26 * 26 * 81
    assert cov[0, 0] == 500.0 and cov[1, 1] == 500.0
    # without expansion only one coarse pass (+ fine)
    o2 = KartoOracle(dict(use_response_expansion=False))
    o2.match(_wall_scan(), (0.0, 0.0, 0.0), [np.zeros((0, 2))], True, True)
    assert o2.dims()["last_passes"] == 2


def test_offsets_table_shape_and_centre():
    o = KartoOracle()
    pts = _wall_scan(50)
    o.build_grid((0.0, 0.0, 0.0), [pts])
    t = o.compute_offsets(pts, (0.0, 0.0, 0.0), 0.0, 0.349, 0.0349)
    assert t.shape == (21, 50)
    stride = o.dims()["stride"]

    def rnd(v):
        return math.floor(v + 0.5) if v >= 0 else math.ceil(v - 0.5)

    exp = np.array([rnd(x * 100) + rnd(y * 100) * stride for x, y in pts])
    # angle index 10 is 0.349 - 10*0.0349 ~ 1e-17 rad: plain cell offsets of the points
    assert (t[10] == exp).all()


def test_grid_is_max_of_stamps():
    o = KartoOracle()
    pts = _wall_scan(300, 2.0, 1.0)
    g = o.build_grid((0.0, 0.0, 0.0), [pts])
    d = o.dims()
    assert g.max() == 100 and g.shape == (d["height"], d["stride"])
    mask = oracle.find_valid_points(pts, 0.0, 0.0)
    k = o.kernel()
    ref = np.zeros_like(g)
    off = 0.0 - 0.5 * (d["roi"] - 1) * 0.01
    for (x, y), ok in zip(pts, mask):
        if not ok:
            continue
        gx = int(math.floor((x - off) * 100 + 0.5)) + d["border"]
        gy = int(math.floor((y - off) * 100 + 0.5)) + d["border"]
        sl = ref[gy - 10:gy + 11, gx - 10:gx + 11]
        np.maximum(sl, k, out=sl)
    assert (g == ref).all()


def test_batch_driver_matches_single(world):
    import scenarios
    b = scenarios.make_batch(world, 6, 360, 2, 11)
    ref = scenarios.oracle_results(None, b, True, True, n_threads=2)
    o = KartoOracle()
    for i in range(6):
        bases = [b["points"][s] for s in b["base_idx"][b["base_ptr"][i]:b["base_ptr"][i + 1]]]
        resp, pose, cov = o.match(b["points"][b["query_scan"][i]], b["query_pose"][i], bases, True, True)
        assert resp == ref[i, 0] and pose == tuple(ref[i, 1:4])
        assert (cov.ravel() == ref[i, 4:]).all()
    assert oracle.lib().ko_probs_collisions(o._h) == 0


@pytest.mark.parametrize("name", ["survey", "blobs", "world"])
def test_raywalk_oracle_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, "raywalk_golden.npz"))
    img = g[f"{name}_img"]
    for an in ("quarter", "coarse"):
        ang, st, ref = g[f"{name}_{an}_angles"], g[f"{name}_{an}_starts"], g[f"{name}_{an}_res"]
        out = oracle.raywalk_sweep_many(img, ang, st)
        assert (out.view(np.uint32) == ref.view(np.uint32)).all()


def test_raywalk_survey_sample_values():
    # SURVEY.md 8c verified sample
    g = np.load(os.path.join(GOLD, "raywalk_golden.npz"))
    out = oracle.raywalk_sweep(g["survey_img"], [0.0, 90.0, 180.0], 100, 100)
    assert tuple(out[0, 2:4]) == (251.0, 100.0) and out[0, 4] == 151.0
    assert tuple(out[1, 2:4]) == (100.0, 1151.0)
    assert tuple(out[2, 2:4]) == (0.0, 100.0)


def test_blockwise_trigger_chain_equals_the_serial_walk():
    """FindValidPoints' trigger chain (0 -> next[0] -> next[next[0]] ...) as k_find_valid finds it: 32 points at
    a time, the visited points of a block by pointer doubling on the lanes (five rounds: reach |= OR of 1 << J over
    the reached lanes, J = J[J]), leaving the block through next[] of its last visited point. A Python model of
    that schedule (the kernel's shuffles / REDUX written as loops) against the serial walk on random chains."""
    rng = np.random.default_rng(0)

    def serial(nxt, n):
        t, out = 0, []
        while t < n:
            out.append(t)
            t = int(nxt[t])
        return out

    def blockwise(nxt, n):
        out, e = [], 0
        while e < n:
            b0 = e & ~31
            nx = [int(nxt[b0 + lane]) if b0 + lane < n else n for lane in range(32)]
            J = [nx[lane] - b0 if nx[lane] < n else 64 for lane in range(32)]
            reach = 1 << (e - b0)
            for _ in range(5):
                add = 0
                for lane in range(32):
                    if (reach >> lane) & 1 and J[lane] < 32:
                        add |= 1 << J[lane]
                JJ = [J[J[lane] & 31] for lane in range(32)]
                reach |= add
                J = [JJ[lane] if J[lane] < 32 else J[lane] for lane in range(32)]
            out.extend(b0 + lane for lane in range(32) if (reach >> lane) & 1)
            e = nx[reach.bit_length() - 1]
        return out

    for _ in range(400):
        n = int(rng.integers(1, 800))
        span = int(rng.integers(1, 40))
        nxt = np.array([min(n, i + 1 + int(rng.integers(0, span))) for i in range(n)])
        assert serial(nxt, n) == blockwise(nxt, n)
