"""Generates tests/golden/matcher_golden.npz from the CPU oracle (oracle/karto_oracle.c) on
seeded synthetic scans (the reference holds no golden vectors for the matcher, so
the build commits its own, plus the reference's test.py geometry as a smoke case).
PARITY UNPINNED w.r.t. the real karto_scanmatcher wheel; these vectors pin the oracle against
regressions and give the GPU tests fixed inputs. Re-run: python tests/golden/make_matcher_golden.py"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import scenarios  # noqa: E402
from oracle import oracle  # noqa: E402
from yag_slam_b200 import synth  # noqa: E402

CASES = {
    # name: (cfg, make_batch kwargs, penalty, do_fine)
    "seq_p360": (None, dict(n_matches=6, n_beams=360, n_base=1, seed=101, perturb=(0.07, 0.03)), True, True),
    "seq_p720": (None, dict(n_matches=6, n_beams=720, n_base=10, seed=102), True, True),
    "seq_nopen": (None, dict(n_matches=4, n_beams=360, n_base=4, seed=103), False, False),
    "loop_deg": (dict(search_size=4.0, resolution=0.05),
                 dict(n_matches=8, n_beams=360, n_base=5, seed=104, perturb=(1.0, 0.2), degenerate_frac=0.3), False, False),
}


def test_py_geometry():
    """reference test.py:23-43: 230 beams of 3.0 m, base at (0,0,0), query at (1.0, 0, 1.57)."""
    ranges = np.full(230, 3.0)
    args = (-1.0, np.deg2rad(0.5), 0.0, 5.0)
    base = oracle.point_readings(ranges, *args, 0.0, 0.0, 0.0)
    query = oracle.point_readings(ranges, *args, 1.0, 0.0, 1.57)
    return base, query, np.array([1.0, 0.0, 1.57])


def main():
    w = synth.make_world()
    out = {}
    for name, (cfg, kw, pen, fine) in CASES.items():
        b = scenarios.make_batch(w, **kw)
        ref = scenarios.oracle_results(cfg, b, pen, fine, n_threads=1)
        for k in ("pool", "starts", "counts", "query_scan", "query_pose", "base_ptr", "base_idx"):
            out[f"{name}_{k}"] = b[k]
        out[f"{name}_ref"] = ref
    base, query, pose = test_py_geometry()
    o = oracle.KartoOracle(dict(range_threshold=20))
    resp, p, cov = o.match(query, pose, [base], True, True)
    out["testpy_base"], out["testpy_query"], out["testpy_pose"] = base, query, pose
    out["testpy_ref"] = np.concatenate([[resp], p, cov.ravel()])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "matcher_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; testpy ->", resp, p)


if __name__ == "__main__":
    main()
