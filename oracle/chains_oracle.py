"""CPU oracle of the loop-closure chain finder (TEST INFRASTRUCTURE ONLY -- never imported by
the product; see oracle/__init__.py).

Restates, on plain arrays, what the reference does per query scan in
  GraphSlam.find_possible_loop_closure_chains      yag_slam/graph_slam.py:274-304
  do_breadth_first_traversal + near_scan_visitor   yag_slam/graph.py:71-98, graph_slam.py:32-39
  RadiusHashSearch.pose_to_key / key_to_pose /
      crude_radius_search                           yag_slam/helpers.py:395-431
  scans_dist_squared / poses_dist_squared           yag_slam/helpers.py:379-386

PINNED: tests/golden/chains_golden.npz holds the output of the reference's own
find_possible_loop_closure_chains (imported unmodified from /root/reference by
tests/golden/make_chains_golden.py); tests/test_chains_cpu.py checks this restatement against it.

Reference quirks kept on purpose:
  * the exact test compares the SQUARED distance with the un-squared loop_search_dist
    (graph_slam.py:291);
  * candidates are walked as consecutive pairs (v1, v2), so the LAST candidate is never
    considered (graph_slam.py:285);
  * an excluded candidate (`continue`, graph_slam.py:287-289) resets the chain and skips the
    gap test of that pair;
  * a trailing partial chain shorter than loop_search_min_chain_size is still returned
    (graph_slam.py:301-302);
  * the hash boxes are keyed with the pose a vertex had when it was hashed (add_vertex or the
    last run_opt), int() truncating toward zero, and the crude test measures from the box's
    corner pose key*res with radius + res (helpers.py:403-431).
"""


def near_linked(pose_xy, adj_ptr, adj_idx, q, distance):
    """Vertices do_breadth_first_traversal(vert, near_scan_visitor) returns (as a set of ids)."""
    distsq = distance ** 2  # graph_slam.py:33
    qx, qy = pose_xy[q]

    def visit(v):  # graph_slam.py:35-37
        return (qx - pose_xy[v][0]) ** 2 + (qy - pose_xy[v][1]) ** 2 < distsq

    to_visit, seen, valid = [q], {q}, []
    while to_visit:  # graph.py:82-95
        v = to_visit.pop()
        if not visit(v):
            continue
        valid.append(v)
        for e in range(adj_ptr[v], adj_ptr[v + 1]):
            u = int(adj_idx[e])
            if u in seen:
                continue
            to_visit.append(u)
            seen.add(u)
    return set(valid)


def crude_candidates(hash_xy, pose_xy, q, radius, res):
    """Sorted vertex ids crude_radius_search(scan.corrected_pose, radius) returns."""
    r2 = (radius + res) ** 2  # helpers.py:423
    qx, qy = pose_xy[q]
    out = []
    for v in range(len(hash_xy)):
        kx, ky = int(hash_xy[v][0] / res), int(hash_xy[v][1] / res)  # helpers.py:405
        bx, by = float(kx) * res, float(ky) * res  # helpers.py:410
        if (bx - qx) ** 2 + (by - qy) ** 2 < r2:  # helpers.py:427, 379-380
            out.append(v)
    return out  # ascending id == sort(key=num), graph_slam.py:282


def find_chains(pose_xy, hash_xy, adj_ptr, adj_idx, q, loop_search_dist, min_chain_size):
    """Chains (lists of vertex ids) for query vertex q."""
    near = near_linked(pose_xy, adj_ptr, adj_idx, q, loop_search_dist)
    cands = crude_candidates(hash_xy, pose_xy, q, loop_search_dist, loop_search_dist)
    qx, qy = pose_xy[q]
    chains, cur = [], []
    for v1, v2 in zip(cands, cands[1:]):  # graph_slam.py:285
        if v1 == q or v1 in near:
            cur = []
            continue
        if (qx - pose_xy[v1][0]) ** 2 + (qy - pose_xy[v1][1]) ** 2 <= loop_search_dist:  # :291
            cur.append(v1)
        if len(cur) >= min_chain_size:
            chains.append(cur)
            cur = []
        if (v2 - v1) > 1:
            cur = []
    if cur:
        chains.append(cur)
    return chains


def find_chains_batch(pose_xy, hash_xy, adj_ptr, adj_idx, queries, loop_search_dist, min_chain_size):
    """CSR form the CUDA path returns: (query_chain_ptr [Q+1], chain_ptr [C+1], members [M])."""
    import numpy as np
    pose = [tuple(map(float, p)) for p in pose_xy]
    hsh = [tuple(map(float, p)) for p in hash_xy]
    ap = [int(v) for v in adj_ptr]
    qcp, cp, mem = [0], [0], []
    for q in queries:
        for ch in find_chains(pose, hsh, ap, adj_idx, int(q), loop_search_dist, min_chain_size):
            mem.extend(ch)
            cp.append(len(mem))
        qcp.append(len(cp) - 1)
    return np.array(qcp, np.int32), np.array(cp, np.int32), np.array(mem, np.int32)
