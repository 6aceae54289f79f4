"""Developer probe under torchrun (one rank per GPU): BASELINE cfg 5 end to end -- the 100,000
relocalisation queries sharded contiguously over the ranks (distributed.match_pool_sharded: no
mid-match exchange, one NCCL all-gather of the 128-B records), then the final map's ray-walk sharded by
start cell (distributed.raytrace_sharded). STRONG scaling: the total work is fixed. Rank 0 re-runs
the whole batch alone and checks the gathered records are byte-identical. Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from yag_slam_b200 import distributed, occupancy, raytracing, synth
    from yag_slam_b200.matcher import ScanMatcherB200
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    w = synth.make_world()
    b = synth.make_relocalisation_batch(w, n, 720, 10, 5)  # seeded: identical on every rank
    cores = len(os.sched_getaffinity(0))
    lanes = 3 if cores // world >= 4 else 2
    m = ScanMatcherB200(None, device=lr, lanes=lanes)
    hpool = torch.from_numpy(b["pool"]).pin_memory()
    args = (hpool, b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"])

    def timed(fn, reps=3):
        best, out = None, None
        for _ in range(reps):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
            best = float(t[0]) if best is None else min(best, float(t[0]))
        return best, out

    sharded = lambda: distributed.match_pool_sharded(m.match_pool, *args, True, True, device=dev)  # noqa: E731
    sharded()  # warm-up
    t_match, full = timed(sharded)
    out = {"probe": "cfg5 relocalisation sharded + all-gather (strong scaling, pinned host pool)", "n_gpus": world,
           "matches": n, "lanes": lanes, "host_cores": cores, "sharded_s": t_match, "matches_per_s": n / t_match}
    # the final map, replicated; ray-walk sharded by start cell
    log = synth.make_scan_log(w, 2000, 720, seed=2)
    g = occupancy.occupancy_grid_from_arrays(log["poses"], log["lasers"], log["ranges"], log["beam_ptr"], 0.05, 12.0,
                                             device=lr)
    img = g.image
    ang = np.arange(1439) * (360.0 / 1439) - 180.0
    free = np.argwhere(img == 255)
    starts = free[np.random.default_rng(0).choice(len(free), 1024, replace=False)][:, ::-1].astype(np.float64)
    trace = lambda i, a, s: raytracing.raytrace_many(i, a, s, device=lr)  # noqa: E731
    rays_fn = lambda: distributed.raytrace_sharded(trace, img, ang, starts, device=dev)  # noqa: E731
    rays_fn()
    t_rays, rays = timed(rays_fn)
    out["raywalk_sharded_ms"] = 1e3 * t_rays
    if rank == 0:
        single = m.match_pool(*args, True, True)
        out["gathered_equals_single_gpu"] = bool(single.tobytes() == full.tobytes())
        out["rays_equal_single_gpu"] = bool(trace(img, ang, starts).tobytes() == rays.tobytes())
        print(json.dumps(out))
    m.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
