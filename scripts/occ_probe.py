"""Developer probe on a GPU box: create_occupancy_grid timing (wall clock, whole call incl. H2D)
vs the CPU oracle on the full cfg-2 log, plus the resident-image ray-walk chain."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from yag_slam_b200 import occupancy, raytracing, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    w = synth.make_world()
    log = synth.make_scan_log(w, n, 720, seed=2)
    args = (log["poses"], log["lasers"], log["ranges"], log["beam_ptr"], 0.05, 12.0)
    g = occupancy.occupancy_grid_from_arrays(*args)
    ts = []
    for _ in range(5):
        g.close()
        t0 = time.perf_counter()
        g = occupancy.occupancy_grid_from_arrays(*args)
        ts.append(time.perf_counter() - t0)
    print("gpu create_occupancy_grid %d scans x 720: %.3f ms (best of 5) %dx%d info %s" %
          (n, 1e3 * min(ts), g.width, g.height, g.info))
    if "--cpu" in sys.argv:
        from oracle import oracle
        t0 = time.perf_counter()
        o = oracle.occupancy_grid(*args)
        t1 = time.perf_counter() - t0
        print("cpu oracle: %.1f ms; image equal: %s" % (1e3 * t1, bool((o["image"] == g.image).all())))
    ang = np.arange(1439) * (360.0 / 1439) - 180.0
    free = np.argwhere(g.image == 255)
    starts = free[np.random.default_rng(0).choice(len(free), 1024, replace=False)][:, ::-1].astype(np.float64)
    raytracing.raytrace_many(g, ang, starts)
    t0 = time.perf_counter()
    out = raytracing.raytrace_many(g, ang, starts)
    print("ray-walk 1024 starts x 1439 angles from the resident image: %.3f ms, mean length %.1f px" %
          (1e3 * (time.perf_counter() - t0), float(out[..., 4].mean())))


if __name__ == "__main__":
    main()
