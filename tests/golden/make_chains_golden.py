"""Generates tests/golden/chains_golden.npz by running the REFERENCE's own
GraphSlam.find_possible_loop_closure_chains (/root/reference/yag_slam/graph_slam.py:274-304,
unmodified, with its RadiusHashSearch and breadth-first traversal) in the build container.
The reference cannot travel to the GPU box, so the vectors are committed.
Re-run:  python tests/golden/make_chains_golden.py

karto_scanmatcher / tiny_tf / sba_cpp are absent here: the value types come from
yag-slam_b200/karto_compat.py and the test shims (tests/shims); none of them takes part in the
chain search itself, which only reads corrected_pose.x/.y, num and the graph edges.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "..", "shims"))
sys.path.insert(0, "/root/reference")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")

import numpy as np  # noqa: E402

from yag_slam_b200 import karto_compat, synth  # noqa: E402

mod = types.ModuleType("karto_scanmatcher")
for n in ("Pose2", "LaserScanConfig", "LocalizedRangeScan", "ScanMatcherConfig", "create_occupancy_grid", "Wrapper"):
    setattr(mod, n, getattr(karto_compat, n))
sys.modules["karto_scanmatcher"] = mod

import yag_slam.graph_slam as gs  # noqa: E402  (reference, unmodified)
import yag_slam.models as models  # noqa: E402
from tiny_tf.tf import Transform  # noqa: E402  (shim)

# name: (n_vertices, step, loop_search_dist, min_chain, pose noise, loop-edge period, stale fraction, seed)
CASES = {
    "default": (700, 0.25, 3, 10, 0.05, 40, 0.0, 11),
    "tight": (500, 0.5, 1.5, 4, 0.10, 25, 0.0, 12),
    "stale_hash": (600, 0.3, 2.5, 6, 0.08, 30, 0.3, 13),
    "no_loop_edges": (400, 0.4, 3.0, 10, 0.05, 0, 0.0, 14),
}


def build_case(n, step, dist, min_chain, noise, loop_every, stale, seed):
    rng = np.random.default_rng(seed)
    path = synth.loop_path(n, step=step)
    path[:, :2] += rng.normal(0, noise, (n, 2)) + np.cumsum(rng.normal(0, noise * 0.05, (n, 2)), axis=0)
    slam = gs.GraphSlam(object(), object(), loop_search_dist=dist, loop_search_min_chain_size=min_chain)
    lp = synth.laser_params(8)
    cov = np.eye(3).tolist()
    edges = []
    per_lap = int(round(70.85 / step))
    for i in range(n):
        s = models.LocalizedRangeScan([1.0] * 8, lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], *path[i])
        s.num = i
        slam.add_vertex(s)  # hashes the vertex at its pose of this moment (graph_slam.py:141-146)
        if i:
            slam.link_scans(slam.graph.vertices[i - 1].obj, s, s.corrected_pose, cov)
            edges.append((i - 1, i))
        if loop_every and i >= per_lap and i % loop_every == 0:
            # a loop-closure link to the closest vertex of the previous lap (link_to_closest_scan_in_chain)
            prev = [v.obj for v in slam.graph.vertices[max(0, i - per_lap - 20):i - per_lap + 20]]
            if prev:
                prev.sort(key=lambda o: gs.scans_dist_squared(o, s))
                slam.link_scans(prev[0], s, s.corrected_pose, cov)
                edges.append((prev[0].num, i))
    hash_xy = np.array([[v.obj.corrected_pose.x, v.obj.corrected_pose.y] for v in slam.graph.vertices])
    if stale > 0:
        # poses move after hashing without a rebuild of the search structure
        for v in slam.graph.vertices:
            if rng.random() < stale:
                p = v.obj.corrected_pose
                v.obj.corrected_pose = Transform.from_position_euler(p.x + rng.normal(0, 0.8), p.y + rng.normal(0, 0.8),
                                                                     0, 0, 0, p.euler[-1])
    pose_xy = np.array([[v.obj.corrected_pose.x, v.obj.corrected_pose.y] for v in slam.graph.vertices])
    queries = np.unique(np.concatenate([np.arange(0, n, 7), np.arange(max(0, n - 40), n)])).astype(np.int32)
    qcp, cp, mem = [0], [0], []
    for q in queries:
        chains = slam.find_possible_loop_closure_chains(slam.graph.vertices[q].obj)
        for ch in chains:
            mem.extend(o.num for o in ch)
            cp.append(len(mem))
        qcp.append(len(cp) - 1)
    return dict(pose_xy=pose_xy, hash_xy=hash_xy, edges=np.array(edges, np.int32).reshape(-1, 2), queries=queries,
                params=np.array([dist, min_chain], np.float64), query_chain_ptr=np.array(qcp, np.int32),
                chain_ptr=np.array(cp, np.int32), members=np.array(mem, np.int32))


def main():
    out = {}
    for name, args in CASES.items():
        d = build_case(*args)
        print(name, "queries", len(d["queries"]), "chains", len(d["chain_ptr"]) - 1, "members", len(d["members"]))
        for k, v in d.items():
            out["%s_%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "chains_golden.npz"), **out)


if __name__ == "__main__":
    main()
