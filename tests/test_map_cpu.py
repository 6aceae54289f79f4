"""CPU tests of the match-against-a-map oracle: its correlation grid against
the golden vectors of the REFERENCE's numba occupancy_grid_map_to_correlation_grid
(tests/golden/make_mapgrid_golden.py), and the match itself on the synthetic world's map."""
import os

import numpy as np
import pytest

from oracle import oracle
from yag_slam_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GRID_CASES = ("world_r05_s05", "world_r05_s10", "blobs_r02_s06")


def load_grid_case(name):
    g = np.load(os.path.join(HERE, "golden", "mapgrid_golden.npz"))
    res, smear = g[name + "_params"]
    return g[name + "_img"], g[name + "_grid"], dict(resolution=float(res), smear_deviation=float(smear), search_size=float(res) * 10)


def map_queries(world, n, P, seed, res=0.05, perturb=(0.15, 0.05)):
    """n query scans at perturbed poses inside the synthetic world + the world's occupancy image."""
    img, off = synth.occupancy_image(world, res)
    rng = np.random.default_rng(seed)
    path = synth.loop_path(n, step=70.0 / n)
    guess = path + np.concatenate([rng.uniform(-perturb[0], perturb[0], (n, 2)), rng.uniform(-perturb[1], perturb[1], (n, 1))], axis=1)
    lp = synth.laser_params(P)
    from yag_slam_b200 import _capi
    pts = [_capi.point_readings(synth.cast_scan(world, path[i], P, rng), lp[0], lp[2], lp[3], lp[5], *guess[i]) for i in range(n)]
    from yag_slam_b200.matcher import pack_pool
    pool, starts, counts = pack_pool(pts)
    return img, off, pool, starts, counts, np.arange(n, dtype=np.int32), guess, path


@pytest.mark.parametrize("name", GRID_CASES)
def test_oracle_map_grid_equals_the_reference_numba_grid(name):
    img, ref, cfg = load_grid_case(name)
    o = oracle.KartoMapOracle(cfg, img, (0.0, 0.0), 0)
    d = o.dims()
    b = d["border"]
    grid = o.grid()
    assert (grid[b:b + img.shape[0], b:b + img.shape[1]] == ref).all()
    assert d["roi"] == max(img.shape) and d["width"] == d["roi"] + 2 * b


def test_oracle_map_match_recovers_the_pose(world):
    img, off, pool, starts, counts, qs, guess, truth = map_queries(world, 6, 360, 4)
    o = oracle.KartoMapOracle(dict(oracle.DEFAULTS, resolution=0.05, search_size=0.5), img, off, 0)
    out = o.match_many(pool, starts, counts, qs, guess, True, True)
    assert (out[:, 0] > 0.3).all()
    assert np.abs(out[:, 1] - truth[:, 0]).max() < 0.11 and np.abs(out[:, 2] - truth[:, 1]).max() < 0.11
    # an empty query keeps its pose (MatchScan early return)
    e = o.match(np.zeros((0, 2)), (1.0, 2.0, 0.3), True, True)
    assert e[0] == 0.0 and tuple(e[1:4]) == (1.0, 2.0, 0.3) and e[4] == 500.0
