"""Generates tests/golden/raywalk_golden.npz by importing the REFERENCE's numba ray-walk
(/root/reference/yag_slam/raytracing.py:63-92) in the build container. The reference cannot
travel to the GPU box, so the vectors are committed. Re-run:  python tests/golden/make_raywalk_golden.py
"""
import os
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np  # noqa: E402
from yag_slam import raytracing  # noqa: E402  (reference, unmodified)


def make_maps():
    maps = {}
    # (1) SURVEY 8c sample: 200x300, wall at x>=250, unknown at y>=150
    m = np.full((200, 300), 255, np.uint8)
    m[:, 250:] = 0
    m[150:, :] = 200
    maps["survey"] = (m, [(100, 100), (10, 20), (240, 140), (100.5, 100.5)])
    # (2) random blobs map with all three classes
    rng = np.random.default_rng(7)
    m = np.full((240, 320), 255, np.uint8)
    for _ in range(25):
        x, y = rng.integers(5, 300), rng.integers(5, 220)
        w, h = rng.integers(2, 18), rng.integers(2, 18)
        m[y:y + h, x:x + w] = 0 if rng.random() < 0.6 else 200
    m[0, :] = 0; m[-1, :] = 0; m[:, 0] = 0; m[:, -1] = 0
    starts = []
    while len(starts) < 6:
        x, y = rng.uniform(2, 317), rng.uniform(2, 237)
        if m[int(round(y)), int(round(x))] == 255:
            starts.append((float(x), float(y)))
    maps["blobs"] = (m, starts)
    # (3) the synthetic world occupancy image (cfg 5 ray-walk shape, cropped)
    from yag_slam_b200 import synth
    img, _ = synth.occupancy_image(synth.make_world(), 0.05)
    maps["world"] = (img, [(420.0, 320.0), (100.25, 500.75), (700.5, 80.5)])
    return maps


def main():
    out = {}
    angle_sets = {
        "quarter": np.arange(-180, 180, 0.25)[:-1][::-1].copy(),   # splicing.py:87,94
        "coarse": np.array([0.0, 10.0, 33.3, 45.0, 90.0, 135.0, 180.0, -90.0, -45.0, 271.5, 359.9]),
    }
    for name, (img, starts) in make_maps().items():
        out[f"{name}_img"] = img
        for an, angles in angle_sets.items():
            if name == "world" and an == "quarter":
                angles = angles[::7].copy()
            res = np.zeros((len(starts), len(angles), 5), np.float32)
            for si, (sx, sy) in enumerate(starts):
                infos = raytracing.run_raytracing_sweep(img, angles, sx, sy)
                for ai, info in enumerate(infos):
                    res[si, ai] = (info.start.x, info.start.y, info.end.x, info.end.y, info.length)
            out[f"{name}_{an}_angles"] = angles
            out[f"{name}_{an}_starts"] = np.array(starts, np.float64)
            out[f"{name}_{an}_res"] = res
    path = os.path.join(os.path.dirname(__file__), "raywalk_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
