"""Developer probe (GPU box): throughput of match-against-a-map -- 4,096 720-beam query scans against the
resident correlation grid of the synthetic world's map (0.05 m/px, loop-matcher search 4.0 m), vs the
CPU oracle (single thread) on a bounded sample; checks bit-equality on the sample."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from oracle import oracle
from yag_slam_b200 import synth
from yag_slam_b200.matcher import DEFAULTS_LOOP, MapMatcherB200
from test_map_cpu import map_queries

world = synth.make_world()
n = 4096
img, off, pool, starts, counts, qs, guess, truth = map_queries(world, n, 720, 9, 0.05, perturb=(1.0, 0.1))
m = MapMatcherB200(DEFAULTS_LOOP, img, off, 0)
for _ in range(3):
    out = m.match_map(pool, starts, counts, qs, guess, False, False)
ts = []
for _ in range(10):
    t0 = time.perf_counter()
    out = m.match_map(pool, starts, counts, qs, guess, False, False)
    ts.append(time.perf_counter() - t0)
ns = 48
o = oracle.KartoMapOracle(DEFAULTS_LOOP, img, off, 0)
t0 = time.perf_counter()
ref = o.match_many(pool, starts, counts, qs[:ns], guess[:ns], False, False)
tcpu = time.perf_counter() - t0
ok = all((out[k][:ns] == ref[:, c]).all() for k, c in (("response", 0), ("x", 1), ("y", 2), ("heading", 3)))
err = np.hypot(out["x"] - truth[:, 0], out["y"] - truth[:, 1])
print(json.dumps({"what": "match against a resident map grid (loop-matcher config, coarse only)", "queries": n, "beams": 720,
                  "map_shape": list(img.shape), "gpu_call_ms_p50": float(np.median(ts) * 1e3),
                  "matches_per_s_e2e": n / float(np.median(ts)), "cpu_port_matches_per_s_1core": ns / tcpu,
                  "cpu_sample": "%d queries" % ns, "bit_exact_on_sample": bool(ok),
                  "median_position_error_m": float(np.median(err)), "work": m.last_work()}))
