"""CPU tests of the loop-closure chain finder: the oracle restatement against
the golden vectors produced by the REFERENCE's own find_possible_loop_closure_chains
(tests/golden/make_chains_golden.py), the host helpers, and the no-GPU failure mode."""
import os

import numpy as np
import pytest

from oracle import chains_oracle as co
from yag_slam_b200 import chains

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ("default", "tight", "stale_hash", "no_loop_edges")


def load_case(name):
    g = np.load(os.path.join(HERE, "golden", "chains_golden.npz"))
    d = {k[len(name) + 1:]: g[k] for k in g.files if k.startswith(name + "_")}
    dist, mc = d["params"]
    d["dist"] = int(dist) if float(dist).is_integer() else float(dist)  # GraphSlam's default is the int 3
    d["min_chain"] = int(mc)
    d["adj_ptr"], d["adj_idx"] = chains.adjacency_csr(len(d["pose_xy"]), d["edges"])
    return d


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_reference_golden(name):
    d = load_case(name)
    a, b, c = co.find_chains_batch(d["pose_xy"], d["hash_xy"], d["adj_ptr"], d["adj_idx"], d["queries"], d["dist"],
                                   d["min_chain"])
    assert (a == d["query_chain_ptr"]).all() and (b == d["chain_ptr"]).all() and (c == d["members"]).all()
    assert len(c) > 100  # the case is not vacuous


def test_adjacency_csr_is_symmetric_and_complete():
    edges = np.array([[0, 1], [1, 2], [0, 2], [3, 1]])
    ptr, idx = chains.adjacency_csr(5, edges)
    nbr = [sorted(idx[ptr[v]:ptr[v + 1]].tolist()) for v in range(5)]
    assert nbr == [[1, 2], [0, 2, 3], [0, 1], [1], []]
    ptr, idx = chains.adjacency_csr(3, np.zeros((0, 2), int))
    assert ptr.tolist() == [0, 0, 0, 0] and len(idx) == 0


def test_oracle_quirks_hand_case():
    # 6 vertices on a line 0.5 m apart, no edges; query is vertex 5; dist 3 -> every box is a candidate.
    pose = [(0.5 * i, 0.0) for i in range(6)]
    ptr = [0] * 7
    # squared distance to v5: 6.25, 4, 2.25, 1, .25 -> "<= 3" keeps 2, 3, 4; the last candidate (5) is never v1;
    # min chain 2 -> [2, 3] is emitted, the partial [4] is kept as a trailing chain
    assert co.find_chains(pose, pose, ptr, [], 5, 3, 2) == [[2, 3], [4]]
    # an edge 4-5 makes 4 "near linked": excluded, and it resets the running chain
    ptr2, idx2 = chains.adjacency_csr(6, [[4, 5]])
    assert co.find_chains(pose, pose, ptr2.tolist(), idx2, 5, 3, 3) == []
    assert co.find_chains(pose, pose, ptr2.tolist(), idx2, 5, 3, 2) == [[2, 3]]


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_chain_finder_fails_loudly_without_gpu():
    d = load_case("tight")
    with pytest.raises(RuntimeError):
        chains.find_chains_batch(d["pose_xy"], d["adj_ptr"], d["adj_idx"], d["queries"], d["dist"], d["min_chain"])
    with pytest.raises(ValueError):
        chains.find_chains_batch(d["pose_xy"], d["adj_ptr"], d["adj_idx"], d["queries"], d["dist"], 0)
