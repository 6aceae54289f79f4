"""Developer probe on a GPU box: BASELINE cfg 5 at full size on ONE GPU -- 100,000 relocalisation
matches (P = 720, 10 base scans, default_config, penalty + fine) through match_pool, with the pool
in pinned host memory (end to end) and resident in HBM, then create_occupancy_grid of the 2,000-scan
log and the 1,439-angle x 1,024-start ray-walk from the resident image. Prints one JSON line.
(The multi-GPU form shards the 100k queries contiguously and all-gathers the records:
yag-slam_b200/distributed.py, bench.py --gpus N.)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    import torch
    from yag_slam_b200 import _capi, occupancy, raytracing, synth
    from yag_slam_b200.matcher import ScanMatcherB200
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    w = synth.make_world()
    t0 = time.perf_counter()
    b = synth.make_relocalisation_batch(w, n, 720, 10, 5)
    gen_s = time.perf_counter() - t0
    m = ScanMatcherB200(None, lanes=lanes)
    res = np.zeros(n, dtype=_capi.RESULT_DTYPE)
    hpool = torch.from_numpy(b["pool"]).pin_memory()
    dpool = torch.from_numpy(b["pool"]).cuda()

    def run(pool):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.match_pool(pool, b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"],
                     True, True, out=res)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    run(dpool)  # warm-up (allocations, first-touch)
    ref = res.copy()
    t_dev = min(run(dpool) for _ in range(3))
    run(hpool)
    t_host = min(run(hpool) for _ in range(3))
    same = bool(res.tobytes() == ref.tobytes())
    e = np.hypot(res["x"] - b["truth"][:, 0], res["y"] - b["truth"][:, 1])
    out = {"probe": "cfg5 relocalisation, 1 GPU", "matches": n, "lanes": lanes, "pool_bytes": int(b["pool"].nbytes),
           "resident_s": t_dev, "resident_matches_per_s": n / t_dev, "e2e_pinned_host_s": t_host,
           "e2e_matches_per_s": n / t_host, "host_equals_resident": same, "status_ok": bool((res["status"] == 0).all()),
           "median_pose_error_m": float(np.median(e)), "generate_s": gen_s}
    m.close()
    # the final map and its ray-walk
    log = synth.make_scan_log(w, 2000, 720, seed=2)
    args = (log["poses"], log["lasers"], log["ranges"], log["beam_ptr"], 0.05, 12.0)
    g = occupancy.occupancy_grid_from_arrays(*args)
    g.close()
    t0 = time.perf_counter()
    g = occupancy.occupancy_grid_from_arrays(*args)
    out["occupancy_grid_ms"] = 1e3 * (time.perf_counter() - t0)
    out["map_wh"] = [int(g.width), int(g.height)]
    ang = np.arange(1439) * (360.0 / 1439) - 180.0
    free = np.argwhere(g.image == 255)
    starts = free[np.random.default_rng(0).choice(len(free), 1024, replace=False)][:, ::-1].astype(np.float64)
    raytracing.raytrace_many(g, ang, starts)
    t0 = time.perf_counter()
    rays = raytracing.raytrace_many(g, ang, starts)
    dt = time.perf_counter() - t0
    out["raywalk_ms"] = 1e3 * dt
    out["rays_per_s"] = rays.shape[0] * rays.shape[1] / dt
    out["mean_ray_px"] = float(rays[..., 4].mean())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
