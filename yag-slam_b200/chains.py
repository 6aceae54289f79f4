"""Batched loop-closure chain finder on the B200: host mirror of the reference's
GraphSlam.find_possible_loop_closure_chains (yag_slam/graph_slam.py:274-304) for MANY query scans
at once, returning the chains in the CSR form `ScanMatcherB200.match_pool` takes as base lists, so
a loop-closure batch (BASELINE cfg 3) is assembled without Python loops.

Host glue only: the search runs in libysm_b200.so (csrc/ysm_chains.cu). No CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _capi

_ERRORS = {_capi.YSM_EINVAL: ValueError, _capi.YSM_EUNSUP: NotImplementedError}


def adjacency_csr(n_vertices, edges):
    """CSR adjacency (both directions, duplicates kept) from an (E, 2) array of (from, to) scan
    numbers -- Vertex.get_adjacent_vertices of every vertex (yag_slam/graph.py:29-36)."""
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    src = np.concatenate([e[:, 0], e[:, 1]])
    dst = np.concatenate([e[:, 1], e[:, 0]])
    order = np.argsort(src, kind="stable")
    ptr = np.zeros(n_vertices + 1, dtype=np.int32)
    np.cumsum(np.bincount(src, minlength=n_vertices), out=ptr[1:])
    return ptr, dst[order].astype(np.int32)


class ChainSet(object):
    """Result of find_chains_batch: CSR arrays + the work counters of the call."""

    def __init__(self, query_chain_ptr, chain_ptr, members, launches, kernel_ms):
        self.query_chain_ptr, self.chain_ptr, self.members = query_chain_ptr, chain_ptr, members
        self.launches, self.kernel_ms = int(launches), float(kernel_ms)

    @property
    def n_chains(self):
        return len(self.chain_ptr) - 1

    def chains_of(self, i):
        """Chains of query i as lists of scan numbers (what the reference returns as scan objects)."""
        return [self.members[self.chain_ptr[c]:self.chain_ptr[c + 1]].tolist()
                for c in range(self.query_chain_ptr[i], self.query_chain_ptr[i + 1])]

    def chain_query(self):
        """[n_chains] index (into the query list) of the query every chain belongs to."""
        return np.repeat(np.arange(len(self.query_chain_ptr) - 1, dtype=np.int32), np.diff(self.query_chain_ptr))


def find_chains_batch(pose_xy, adj_ptr, adj_idx, query_vertices, loop_search_dist=3, loop_search_min_chain_size=10,
                      hash_xy=None, device=0, stream=0):
    """Chains for every query vertex. pose_xy: (n, 2) current corrected poses indexed by scan number;
    hash_xy: poses the vertices had when RadiusHashSearch hashed them (default: pose_xy);
    adj_ptr/adj_idx: adjacency_csr(...) of the pose graph. Defaults are GraphSlam's
    (yag_slam/graph_slam.py:48-49)."""
    pose = np.ascontiguousarray(pose_xy, dtype=np.float64).reshape(-1, 2)
    hsh = None if hash_xy is None else np.ascontiguousarray(hash_xy, dtype=np.float64).reshape(-1, 2)
    if hsh is not None and len(hsh) != len(pose):
        raise ValueError("hash_xy / pose_xy size mismatch")
    ap = np.ascontiguousarray(adj_ptr, dtype=np.int32)
    ai = np.ascontiguousarray(adj_idx, dtype=np.int32)
    qv = np.ascontiguousarray(query_vertices, dtype=np.int32)
    if len(ap) != len(pose) + 1:
        raise ValueError("adj_ptr must have n_vertices + 1 entries")
    q = _capi.YsmChainQuery()
    q.n_vertices, q.n_queries = len(pose), len(qv)
    q.pose_xy = pose.ctypes.data if len(pose) else None
    q.hash_xy = hsh.ctypes.data if hsh is not None else None
    q.adj_ptr = ap.ctypes.data
    q.adj_idx = ai.ctypes.data if len(ai) else None
    q.query_vertex = qv.ctypes.data if len(qv) else None
    # the two thresholds are evaluated exactly as the reference's Python does
    # (helpers.py:423: (radius + self.res)**2 with radius == res == loop_search_dist; graph_slam.py:33)
    q.loop_search_dist = float(loop_search_dist)
    q.crude_r2 = float((loop_search_dist + loop_search_dist) ** 2)
    q.near_dist_sq = float(loop_search_dist ** 2)
    q.min_chain_size = int(loop_search_min_chain_size)
    L = _capi.lib()
    h = C.c_void_p()
    rc = L.ysm_chains_find(C.byref(q), int(device), C.c_void_p(int(stream)), C.byref(h))
    if rc != _capi.YSM_OK:
        raise _ERRORS.get(rc, RuntimeError)(L.ysm_chains_last_error().decode("utf-8", "replace"))
    try:
        nc, nm, nl, ms = C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
        L.ysm_chains_get_counts(h, C.byref(nc), C.byref(nm), C.byref(nl), C.byref(ms))
        qcp = np.zeros(len(qv) + 1, dtype=np.int32)
        cp = np.zeros(nc.value + 1, dtype=np.int32)
        mem = np.zeros(max(nm.value, 1), dtype=np.int32)
        L.ysm_chains_copy(h, qcp.ctypes.data, cp.ctypes.data, mem.ctypes.data)
    finally:
        L.ysm_chains_destroy(h)
    return ChainSet(qcp, cp, mem[:nm.value], nl.value, ms.value)


def graph_arrays(graph_slam):
    """(pose_xy, adj_ptr, adj_idx) of a reference GraphSlam object (yag_slam/graph_slam.py:42-71):
    vertices are indexed by scan number, edges carry .source / .target vertices."""
    verts = graph_slam.graph.vertices
    pose = np.array([[v.obj.corrected_pose.x, v.obj.corrected_pose.y] for v in verts], dtype=np.float64).reshape(-1, 2)
    edges = np.array([[e.source.obj.num, e.target.obj.num] for e in graph_slam.graph.edges], dtype=np.int64).reshape(-1, 2)
    ptr, idx = adjacency_csr(len(verts), edges)
    return pose, ptr, idx


def loop_closure_batch(chains, query_vertices):
    """match_pool arguments for "every query against each of its chains" (graph_slam.py:216-220):
    returns (query_scan [n_chains], base_ptr [n_chains+1], base_idx [members]) over a pool whose
    scan i is vertex i."""
    qv = np.asarray(query_vertices, dtype=np.int32)
    return qv[chains.chain_query()], chains.chain_ptr.copy(), chains.members.copy()
