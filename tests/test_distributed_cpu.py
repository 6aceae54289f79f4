"""world_size-2 gloo test of the multi-GPU host logic (shard -> match -> all-gather): the
gathered array must equal the single-process result. The per-rank matcher is a deterministic
stand-in (host logic only; the CUDA path is covered by the gpu tests)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from yag_slam_b200 import _capi, distributed


def _fake_match(pool, starts, counts, query_scan, query_pose, base_ptr, base_idx, penalty, do_fine):
    out = np.zeros(len(query_scan), dtype=_capi.RESULT_DTYPE)
    for i in range(len(query_scan)):
        nb = base_ptr[i + 1] - base_ptr[i]
        out["response"][i] = 0.001 * query_scan[i] + 0.01 * nb + (0.5 if penalty else 0.0)
        out["x"][i], out["y"][i], out["heading"][i] = query_pose[i]
        out["cov"][i] = np.arange(9) + float(base_idx[base_ptr[i]:base_ptr[i + 1]].sum())
        out["n_passes"][i] = 2 if do_fine else 1
    return out


def _fake_trace(img, angles, starts):
    a = np.asarray(angles, dtype=np.float32)
    s = np.asarray(starts, dtype=np.float32).reshape(-1, 2)
    out = np.zeros((len(s), len(a), 5), dtype=np.float32)
    out[:, :, 0], out[:, :, 1] = s[:, None, 0], s[:, None, 1]
    out[:, :, 2] = s[:, None, 0] + np.cos(np.deg2rad(a))[None, :]
    out[:, :, 3] = s[:, None, 1] + np.sin(np.deg2rad(a))[None, :]
    out[:, :, 4] = 1.0 + s[:, None, 0] * 0.01
    return out


def _starts(n=7):
    return np.stack([np.arange(n) * 3.0 + 1, np.arange(n) * 2.0 + 5], axis=1)


def _problem(n=11):
    rng = np.random.default_rng(0)
    nb = rng.integers(1, 4, n)
    base_ptr = np.concatenate([[0], np.cumsum(nb)]).astype(np.int32)
    base_idx = rng.integers(0, 50, base_ptr[-1]).astype(np.int32)
    return (np.zeros((4, 2)), np.zeros(50, np.int32), np.zeros(50, np.int32), np.arange(n, dtype=np.int32),
            rng.normal(size=(n, 3)), base_ptr, base_idx)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        args = _problem()
        full = distributed.match_pool_sharded(_fake_match, *args, penalty=True, do_fine=True)
        rays = distributed.raytrace_sharded(_fake_trace, None, np.arange(0, 360, 45.0), _starts())
        q.put((rank, full.tobytes() + rays.tobytes()))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 8, 100000):
        for w in (1, 2, 3, 8):
            r = [distributed.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_two_rank_gather_equals_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _fake_match(*_problem(), True, True).tobytes() + _fake_trace(None, np.arange(0, 360, 45.0), _starts()).tobytes()
    assert got[0] == ref and got[1] == ref
