"""GPU parity of match-against-a-map (ysm_create_map + ysm_match_batch on the resident grid) through
the C ABI: the grid against the golden vectors of the reference's numba map-to-grid function, the
matches bit-exact against the oracle (incl. queries hanging over the map edge, empty queries,
batches of several waves)."""
import numpy as np
import pytest

from oracle import oracle
from yag_slam_b200.matcher import DEFAULTS, MapMatcherB200

from test_map_cpu import GRID_CASES, load_grid_case, map_queries

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GRID_CASES)
def test_map_grid_equals_the_reference_numba_grid(name):
    img, ref, cfg = load_grid_case(name)
    m = MapMatcherB200(cfg, img, (0.0, 0.0), 0)
    assert (m.correlation_grid() == ref).all()
    o = oracle.KartoMapOracle(cfg, img, (0.0, 0.0), 0)
    assert (m.debug_grid(0) == o.grid()).all()  # border included
    m.close()


@pytest.mark.parametrize("res,search,P,n,fine", [(0.05, 0.5, 360, 40, True), (0.05, 2.0, 720, 12, False),
                                                 (0.025, 0.5, 360, 10, True)])
def test_map_matches_vs_oracle(world, res, search, P, n, fine):
    cfg = dict(DEFAULTS, resolution=res, search_size=search, smear_deviation=2 * res)
    img, off, pool, starts, counts, qs, guess, truth = map_queries(world, n, P, 11, res)
    m = MapMatcherB200(cfg, img, off, 0)
    out = m.match_map(pool, starts, counts, qs, guess, True, fine)
    ref = oracle.KartoMapOracle(cfg, img, off, 0).match_many(pool, starts, counts, qs, guess, True, fine)
    for k, c in (("response", 0), ("x", 1), ("y", 2), ("heading", 3)):
        assert (out[k] == ref[:, c]).all(), k
    assert np.allclose(out["cov"], ref[:, 4:], rtol=1e-5, atol=0)
    assert (out["response"] > 0.25).mean() > 0.8
    # single-query calls take the small-batch path: same records
    one = m.match_map(pool, starts, counts, qs[:1], guess[:1], True, fine)
    assert one[0].tobytes() == out[0].tobytes()
    m.close()


def test_edge_overhang_empty_query_and_many_waves(world):
    cfg = dict(DEFAULTS, resolution=0.05, search_size=1.0)
    img, off, pool, starts, counts, qs, guess, truth = map_queries(world, 30, 180, 5)
    crop = img[100:420, 150:700].copy()  # most scans now see walls that are outside the map
    coff = (off[0] + 150 * 0.05, off[1] + 100 * 0.05)
    # one empty scan appended to the pool
    starts2 = np.concatenate([starts, [len(pool)]]).astype(np.int32)
    counts2 = np.concatenate([counts, [0]]).astype(np.int32)
    reps = 300  # 9,030 queries: three waves of the map handle
    q_all = np.concatenate([np.tile(qs, reps), [len(starts)] * 30]).astype(np.int32)
    g_all = np.concatenate([np.tile(guess, (reps, 1)), guess])
    m = MapMatcherB200(cfg, crop, coff, 0)
    out = m.match_map(pool, starts2, counts2, q_all, g_all, False, False)
    o = oracle.KartoMapOracle(cfg, crop, coff, 0)
    ref = o.match_many(pool, starts2, counts2, q_all[:30], g_all[:30], False, False)
    for k, c in (("response", 0), ("x", 1), ("y", 2), ("heading", 3)):
        assert (out[k][:30] == ref[:, c]).all(), k
        assert (out[k][:9000].reshape(reps, 30) == out[k][:30]).all(), k  # every repetition identical
    assert (out["response"][9000:] == 0).all() and (out["x"][9000:] == guess[:, 0]).all()
    assert (out["cov"][9000:, 0] == 500.0).all()
    m.close()


def test_map_argument_errors():
    with pytest.raises(ValueError):
        MapMatcherB200(None, np.zeros((0, 4), np.uint8), (0, 0))
    with pytest.raises((ValueError, RuntimeError)):
        MapMatcherB200(dict(resolution=0.05, smear_deviation=1.0), np.zeros((8, 8), np.uint8), (0, 0))
