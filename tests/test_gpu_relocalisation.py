"""BASELINE cfg 5 at full size: 100,000 independent (query, 10-scan base set) relocalisation
matches (P = 720, default_config, pose perturbation U(+-0.2 m, +-0.15 rad)) through the
C ABI. The oracle cannot finish 100k matches in seconds, so the test checks size-independent
properties (shard independence = what the multi-GPU partition relies on, duplicate queries,
recovery of the true pose) plus a bit-exact oracle comparison on a strided sample."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N, P, NB = 100_000, 720, 10


def _sub(b, idx):
    idx = np.asarray(idx)
    return dict(b, query_scan=b["query_scan"][idx], query_pose=b["query_pose"][idx],
                base_ptr=(np.arange(len(idx) + 1) * NB).astype(np.int32),
                base_idx=b["base_idx"].reshape(-1, NB)[idx].reshape(-1))


def _run(m, b):
    return m.match_pool(b["pool"], b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"],
                        b["base_idx"], True, True)


def test_fullsize_relocalisation_batch_properties(world):
    import scenarios
    from test_gpu_parity import _assert_parity
    from yag_slam_b200 import synth
    from yag_slam_b200.matcher import ScanMatcherB200
    b = synth.make_relocalisation_batch(world, N, P, NB, 5)
    m = ScanMatcherB200(None, lanes=3)
    out = _run(m, b)
    assert m.launch_count() > 0 and (out["status"] == 0).all()
    # (a) shard independence: a contiguous 1/8 shard (what rank 5 of 8 would match) and a strided
    # subset give the records of the full batch, bit for bit
    lo, hi = 5 * N // 8, 6 * N // 8
    assert _run(m, _sub(b, np.arange(lo, hi))).tobytes() == out[lo:hi].tobytes()
    idx = np.arange(17, N, 1543)
    sub = _run(m, _sub(b, idx))
    assert sub.tobytes() == out[idx].tobytes()
    # (b) the same sample against the oracle: response / pose bit-exact, covariance <= 1e-5 rel
    _assert_parity(sub, scenarios.oracle_results(None, _sub(b, idx), True, True), "cfg5 sample")
    # (c) duplicates: a query listed twice (anywhere in the batch) gives the same record
    dup = np.concatenate([idx, idx[::-1]])
    d = _run(m, _sub(b, dup))
    assert d[:len(idx)].tobytes() == sub.tobytes() and d[len(idx):].tobytes() == sub[::-1].tobytes()
    # (d) relocalisation works: the matched pose is closer to the truth than the guess was
    e_out = np.hypot(out["x"] - b["truth"][:, 0], out["y"] - b["truth"][:, 1])
    e_in = np.hypot(b["query_pose"][:, 0] - b["truth"][:, 0], b["query_pose"][:, 1] - b["truth"][:, 1])
    assert np.median(e_out) < 0.03 and (e_out < e_in).mean() > 0.9
    # responses are penalised averages of byte sums: within (0, 1]
    assert (out["response"] > 0).all() and (out["response"] <= 1.0).all()
    m.close()
