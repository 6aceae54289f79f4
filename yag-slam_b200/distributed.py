"""Multi-GPU sharding of independent match queries.

Every MatchScan is independent given (query, base set, params): a batch is partitioned into
contiguous ranges, one per rank (one process per GPU), nothing is exchanged mid-match, and
one all-gather of the fixed 128-byte result records follows (NCCL over NVLink on GPUs; gloo
in the CPU tests of this host logic)."""
import numpy as np

from . import _capi


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of rank `rank` out of `world` for n items (sizes differ by <= 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def slice_batch(query_scan, query_pose, base_ptr, base_idx, lo, hi):
    """CSR slice of the match descriptors for matches [lo, hi)."""
    base_ptr = np.asarray(base_ptr, dtype=np.int32)
    b0, b1 = int(base_ptr[lo]), int(base_ptr[hi])
    return (np.asarray(query_scan, dtype=np.int32)[lo:hi],
            np.asarray(query_pose, dtype=np.float64).reshape(-1, 3)[lo:hi],
            (base_ptr[lo:hi + 1] - b0).astype(np.int32),
            np.asarray(base_idx, dtype=np.int32)[b0:b1])


def all_gather_results(local, n_total, group=None, device=None):
    """All-gather per-rank result records (structured array, 128 B each) into the full batch
    order. Uses torch.distributed (backend of the default group: nccl or gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-n_total // world) if n_total else 0  # padded shard size
    buf = np.zeros((max(per, 1), 16), dtype=np.float64)
    lo, hi = shard_range(n_total, rank, world)
    assert len(local) == hi - lo
    if len(local):
        buf[:hi - lo] = np.ascontiguousarray(local).view(np.float64).reshape(-1, 16)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device, non_blocking=False)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out.view(-1, 16), t, group=group)
    out = out.cpu().numpy()
    full = np.zeros(n_total, dtype=_capi.RESULT_DTYPE)
    for r in range(world):
        l, h = shard_range(n_total, r, world)
        if h > l:
            full[l:h] = np.ascontiguousarray(out[r, :h - l]).view(_capi.RESULT_DTYPE).reshape(-1)
    return full


def match_pool_sharded(match_fn, pool_xy, scan_start, scan_count, query_scan, query_pose, base_ptr, base_idx,
                       penalty=True, do_fine=False, group=None, device=None):
    """Each rank matches its contiguous shard with `match_fn` (ScanMatcherB200.match_pool or any
    callable with that signature), then all ranks receive every result."""
    import torch.distributed as dist

    n = len(query_scan)
    if not (dist.is_available() and dist.is_initialized()):
        return match_fn(pool_xy, scan_start, scan_count, query_scan, query_pose, base_ptr, base_idx, penalty, do_fine)
    lo, hi = shard_range(n, dist.get_rank(group), dist.get_world_size(group))
    qs, qp, bp, bi = slice_batch(query_scan, query_pose, base_ptr, base_idx, lo, hi)
    if hi > lo:
        local = match_fn(pool_xy, scan_start, scan_count, qs, qp, bp, bi, penalty, do_fine)
    else:
        local = np.zeros(0, dtype=_capi.RESULT_DTYPE)
    return all_gather_results(local, n, group=group, device=device)


def raytrace_sharded(trace_fn, img, angles_deg, starts_xy, group=None, device=None):
    """Ray-walk sweeps from many start cells, sharded by start cell with the map replicated on every
    rank: each rank runs `trace_fn(img, angles, starts[lo:hi])`
    (raytracing.raytrace_many or any callable of that shape -> (n, n_angles, 5) float32), then one
    all-gather returns the (n_starts, n_angles, 5) array to every rank."""
    import torch
    import torch.distributed as dist

    starts = np.ascontiguousarray(starts_xy, dtype=np.float64).reshape(-1, 2)
    na, n = len(angles_deg), len(starts)
    if not (dist.is_available() and dist.is_initialized()):
        return trace_fn(img, angles_deg, starts)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(n, rank, world)
    per = -(-n // world) if n else 0
    buf = np.zeros((max(per, 1), na, 5), dtype=np.float32)
    if hi > lo:
        buf[:hi - lo] = trace_fn(img, angles_deg, starts[lo:hi])
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out.view(-1, na, 5), t, group=group)
    out = out.cpu().numpy()
    full = np.zeros((n, na, 5), dtype=np.float32)
    for r in range(world):
        l, h = shard_range(n, r, world)
        full[l:h] = out[r, :h - l]
    return full
