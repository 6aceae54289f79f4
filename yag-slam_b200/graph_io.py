"""Offline re-matching of a saved yag_slam graph, straight from its wire format.

The reference checkpoints a GraphSlam as zlib(msgpack(dict)) (yag_slam/graph_slam.py:77-105,
serde keys yag_slam/serde.py:56-95). This module
  * reads that blob into flat arrays (no reference code, no per-scan Python objects),
  * turns "every scan against its scan_buffer_len running scans" (GraphSlam.process_scan,
    graph_slam.py:306-339) into ONE match_pool batch on the B200,
  * appends the 128-byte result records to the same blob under the extra key `match_results`;
    GraphSlam.deserialize (graph_slam.py:108-130) reads its keys by name, so a file with a results
    log still loads in the unmodified reference.
Host glue only: the matching itself is ScanMatcherB200.match_pool (CUDA; no CPU fallback).
"""
import math
import zlib

import msgpack
import numpy as np

from . import _capi
from .matcher import pack_pool

RESULTS_KEY = "match_results"


def _yaw(t):
    """euler[-1] of a serialised Transform (serde.py:94: x, y, z, qx, qy, qz, qw)."""
    return math.atan2(2.0 * (t["qw"] * t["qz"] + t["qx"] * t["qy"]), 1.0 - 2.0 * (t["qy"] * t["qy"] + t["qz"] * t["qz"]))


def _compose(a, b):
    """a + b of tiny_tf Transforms restricted to SE(2): pose b expressed in frame a."""
    c, s = math.cos(a[2]), math.sin(a[2])
    return (a[0] + c * b[0] - s * b[1], a[1] + s * b[0] + c * b[1], a[2] + b[2])


def _relative(a, b):
    """a - b: pose a expressed in frame b (b^-1 o a)."""
    c, s = math.cos(-b[2]), math.sin(-b[2])
    dx, dy = a[0] - b[0], a[1] - b[1]
    return (c * dx - s * dy, s * dx + c * dy, a[2] - b[2])


class SavedGraph(object):
    """Flat view of a GraphSlam checkpoint."""

    def __init__(self, d):
        self.raw = d
        scans = d["scans"]
        n = len(scans)
        self.n = n
        self.num = np.array([s["num"] for s in scans], dtype=np.int32)
        self.ranges = [np.asarray(s["ranges"], dtype=np.float64) for s in scans]
        # min_angle, angle_increment, min_range, range_threshold (what LocalizedRangeScan::Update uses)
        self.laser = np.array([[s["min_angle"], s["angle_increment"], s["min_range"], s["range_threshold"]] for s in scans],
                              dtype=np.float64).reshape(n, 4)
        self.max_angle = np.array([s["max_angle"] for s in scans], dtype=np.float64)
        self.max_range = np.array([s["max_range"] for s in scans], dtype=np.float64)
        self.odom = np.array([[s["odom_pose"]["x"], s["odom_pose"]["y"], _yaw(s["odom_pose"])] for s in scans],
                             dtype=np.float64).reshape(n, 3)
        self.corrected = np.array([[s["corrected_pose"]["x"], s["corrected_pose"]["y"], _yaw(s["corrected_pose"])]
                                   for s in scans], dtype=np.float64).reshape(n, 3)
        self.edges = np.array([[e[0], e[1]] for e in d["edges"]], dtype=np.int32).reshape(-1, 2)
        self.running_scans = list(d.get("running_scans", []))
        self.scan_buffer_len = int(d.get("scan_buffer_len", 10))
        self.loop_search_dist = d.get("loop_search_dist", 3)
        self.loop_search_min_chain_size = int(d.get("loop_search_min_chain_size", 10))
        self.seq_matcher_config = {k: v for k, v in (d.get("seq_matcher_config") or {}).items() if k != "___name"}
        lc = d.get("loop_matcher_config")
        self.loop_matcher_config = {k: v for k, v in lc.items() if k != "___name"} if lc else None

    @property
    def results(self):
        """The results log of the file as a structured array (None if the file has none)."""
        r = self.raw.get(RESULTS_KEY)
        if not r:
            return None
        out = np.frombuffer(r["records"], dtype=_capi.RESULT_DTYPE).copy()
        return dict(records=out, query=np.asarray(r["query"], np.int32), base_ptr=np.asarray(r["base_ptr"], np.int32),
                    base_idx=np.asarray(r["base_idx"], np.int32), guess=r["guess"], penalty=r["penalty"],
                    do_fine=r["do_fine"])


def loads(blob):
    """GraphSlam.unbinarize's decoding (graph_slam.py:94-96) without building objects."""
    return SavedGraph(msgpack.unpackb(zlib.decompress(blob), raw=False, strict_map_key=False))


def load(path):
    with open(path, "rb") as f:
        return loads(f.read())


def rematch_batch(g, guess="odom"):
    """The (pool, descriptors) of "scan k against the scan_buffer_len scans before it" for every k >= 1.

    guess="odom": the initial pose process_scan uses, last.corrected + (query.odom - last.odom)
    (graph_slam.py:320-324); guess="stored": the stored corrected pose (re-check of a converged graph).
    Base scans sit at their stored corrected poses. Returns the match_pool arguments as a dict."""
    pts = [_capi.point_readings(g.ranges[i], g.laser[i, 0], g.laser[i, 1], g.laser[i, 2], g.laser[i, 3], *g.corrected[i])
           for i in range(g.n)]
    query_pose = np.zeros((max(g.n - 1, 0), 3))
    qpts = []
    for k in range(1, g.n):
        if guess == "odom":
            p = _compose(tuple(g.corrected[k - 1]), _relative(tuple(g.odom[k]), tuple(g.odom[k - 1])))
        elif guess == "stored":
            p = tuple(g.corrected[k])
        else:
            raise ValueError("guess must be 'odom' or 'stored'")
        query_pose[k - 1] = p
        qpts.append(_capi.point_readings(g.ranges[k], g.laser[k, 0], g.laser[k, 1], g.laser[k, 2], g.laser[k, 3], *p))
    pool, starts, counts = pack_pool(pts + qpts)  # scans 0..n-1 at stored poses, then the queries at their guess
    L = g.scan_buffer_len
    base_ptr, base_idx = [0], []
    for k in range(1, g.n):
        base_idx.extend(range(max(0, k - L), k))
        base_ptr.append(len(base_idx))
    return dict(pool=pool, starts=starts, counts=counts, query_scan=np.arange(g.n, 2 * g.n - 1, dtype=np.int32),
                query_pose=query_pose, base_ptr=np.array(base_ptr, np.int32), base_idx=np.array(base_idx, np.int32),
                query=np.arange(1, g.n, dtype=np.int32))


def rematch(g, matcher, guess="odom", penalty=True, do_fine=True):
    """Runs the batch on `matcher` (a ScanMatcherB200 built from g.seq_matcher_config) and returns
    (records, batch)."""
    b = rematch_batch(g, guess)
    rec = matcher.match_pool(b["pool"], b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"],
                             b["base_idx"], penalty, do_fine)
    return rec, b


def dumps_with_results(g, records, batch, guess, penalty, do_fine):
    """The checkpoint blob with the results log added (same zlib(msgpack) framing)."""
    d = dict(g.raw)
    d[RESULTS_KEY] = dict(records=np.ascontiguousarray(records).tobytes(), query=batch["query"].tolist(),
                          base_ptr=batch["base_ptr"].tolist(), base_idx=batch["base_idx"].tolist(), guess=guess,
                          penalty=bool(penalty), do_fine=bool(do_fine), record_bytes=_capi.RESULT_DTYPE.itemsize,
                          record_fields=list(_capi.RESULT_DTYPE.names))
    return zlib.compress(msgpack.packb(d, use_bin_type=True))
