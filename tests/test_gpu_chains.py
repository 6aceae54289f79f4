"""GPU parity of the batched loop-closure chain finder (csrc/ysm_chains.cu) through the C ABI:
against the golden vectors of the REFERENCE's find_possible_loop_closure_chains, against the
oracle on larger random pose graphs, and chained into a loop-closure match batch."""
import numpy as np
import pytest

from oracle import chains_oracle as co
from oracle import oracle
from yag_slam_b200 import chains, synth
from yag_slam_b200.matcher import DEFAULTS_LOOP, ScanMatcherB200

from test_chains_cpu import CASES, load_case

pytestmark = pytest.mark.gpu


def _same(cs, ref):
    a, b, c = ref
    return (cs.query_chain_ptr == a).all() and (cs.chain_ptr == b).all() and (cs.members == c).all()


@pytest.mark.parametrize("name", CASES)
def test_reference_golden(name):
    d = load_case(name)
    cs = chains.find_chains_batch(d["pose_xy"], d["adj_ptr"], d["adj_idx"], d["queries"], d["dist"], d["min_chain"],
                                  hash_xy=d["hash_xy"])
    assert _same(cs, (d["query_chain_ptr"], d["chain_ptr"], d["members"]))
    assert cs.launches >= 1
    # chains_of mirrors the reference's list-of-lists return value
    q0 = int(np.argmax(np.diff(d["query_chain_ptr"])))
    assert cs.chains_of(q0) == [d["members"][d["chain_ptr"][c]:d["chain_ptr"][c + 1]].tolist()
                                for c in range(d["query_chain_ptr"][q0], d["query_chain_ptr"][q0 + 1])]


@pytest.mark.parametrize("n,dist,mc,seed", [(3000, 3, 10, 1), (2500, 2.0, 5, 2), (1200, 0.7, 1, 3)])
def test_random_pose_graphs_vs_oracle(n, dist, mc, seed):
    rng = np.random.default_rng(seed)
    path = synth.loop_path(n, step=0.2)[:, :2] + rng.normal(0, 0.15, (n, 2))
    path[::97] -= 40.0  # far outliers with negative coordinates (int() truncation toward zero)
    seq = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1)
    extra = rng.integers(0, n, (n // 10, 2))
    ptr, idx = chains.adjacency_csr(n, np.concatenate([seq, extra]))
    hsh = path + rng.normal(0, 0.3, (n, 2)) * (rng.random((n, 1)) < 0.2)
    queries = rng.integers(0, n, 300).astype(np.int32)
    cs = chains.find_chains_batch(path, ptr, idx, queries, dist, mc, hash_xy=hsh)
    assert _same(cs, co.find_chains_batch(path, hsh, ptr, idx, queries, dist, mc))
    assert cs.n_chains > 0


def test_edge_cases():
    pose = np.array([[0.0, 0.0]])
    cs = chains.find_chains_batch(pose, [0, 0], [], [0], 3, 10)
    assert cs.n_chains == 0 and cs.query_chain_ptr.tolist() == [0, 0]
    cs = chains.find_chains_batch(pose, [0, 0], [], [], 3, 10)
    assert cs.n_chains == 0 and cs.query_chain_ptr.tolist() == [0]
    with pytest.raises(ValueError):
        chains.find_chains_batch(pose, [0, 0], [], [1], 3, 10)  # query out of range
    with pytest.raises(ValueError):
        chains.find_chains_batch(pose, [0, 0], [], [0], 3, 0)


def test_chains_feed_a_loop_closure_match_batch(world):
    """cfg-3 flow: chains found on the device become the CSR base lists of match_pool (loop matcher,
    penalty False / do_fine False, reference graph_slam.py:220) -- results equal the oracle's."""
    n, P = 420, 180
    rng = np.random.default_rng(5)
    path = synth.loop_path(n, step=0.25)
    path[:, :2] += rng.normal(0, 0.03, (n, 2))
    seq = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1)
    ptr, idx = chains.adjacency_csr(n, seq)
    queries = np.arange(300, 420, 6).astype(np.int32)  # second lap: first-lap scans are loop candidates
    cs = chains.find_chains_batch(path[:, :2], ptr, idx, queries, 3, 10)
    assert cs.n_chains >= len(queries)
    qscan, base_ptr, base_idx = chains.loop_closure_batch(cs, queries)
    pts = [synth.scan_points(world, path[i], P, rng) for i in range(n)]
    from yag_slam_b200.matcher import pack_pool
    pool, starts, counts = pack_pool(pts)
    m = ScanMatcherB200(DEFAULTS_LOOP, max_slots=64)
    out = m.match_pool(pool, starts, counts, qscan, path[qscan], base_ptr, base_idx, False, False)
    ref = oracle.match_batch(DEFAULTS_LOOP, pool, starts, counts, qscan, path[qscan], base_ptr, base_idx, False, False)
    for k, c in (("response", 0), ("x", 1), ("y", 2), ("heading", 3)):
        assert (out[k] == ref[:, c]).all(), k
    assert (out["response"] > 0.3).mean() > 0.5
    m.close()
