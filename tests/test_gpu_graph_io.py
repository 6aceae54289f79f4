"""GPU test of the offline re-matching path: a checkpoint written by the
reference's GraphSlam.to_file (tests/golden/graph_checkpoint.bin) -> flat arrays -> ONE match_pool
batch on the B200 -> results log; bit-exact against the oracle on the same batch."""
import os

import numpy as np
import pytest

from oracle import oracle
from yag_slam_b200 import graph_io
from yag_slam_b200.matcher import ScanMatcherB200

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("guess", ["odom", "stored"])
def test_rematch_saved_graph(guess):
    g = graph_io.load(os.path.join(HERE, "golden", "graph_checkpoint.bin"))
    assert g.n == 12 and g.scan_buffer_len == 5
    m = ScanMatcherB200(g.seq_matcher_config, max_slots=16)
    rec, b = graph_io.rematch(g, m, guess, True, True)
    ref = oracle.match_batch(g.seq_matcher_config, b["pool"], b["starts"], b["counts"], b["query_scan"], b["query_pose"],
                             b["base_ptr"], b["base_idx"], True, True)
    for k, c in (("response", 0), ("x", 1), ("y", 2), ("heading", 3)):
        assert (rec[k] == ref[:, c]).all(), k
    assert np.allclose(rec["cov"], ref[:, 4:], rtol=1e-5, atol=0)
    # the re-matched poses agree with the corrected poses the mapping run stored (same inputs, same matcher semantics)
    assert np.abs(rec["x"] - g.corrected[1:, 0]).max() < 0.03 and np.abs(rec["y"] - g.corrected[1:, 1]).max() < 0.03
    assert (rec["response"] > 0.5).all()
    blob = graph_io.dumps_with_results(g, rec, b, guess, True, True)
    r = graph_io.loads(blob).results
    assert (r["records"].view(np.uint8) == rec.view(np.uint8)).all() and r["guess"] == guess
    m.close()
