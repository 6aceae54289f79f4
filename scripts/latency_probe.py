"""Developer probe: single-match latency breakdown (YSM_TRACE=1 prints C-side phases)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from yag_slam_b200 import synth, karto_compat as kc
from yag_slam_b200.matcher import pack_pool

world = synth.make_world()
P, nb = int(sys.argv[1]) if len(sys.argv) > 1 else 360, int(sys.argv[2]) if len(sys.argv) > 2 else 1
lp = synth.laser_params(P)
rng = np.random.default_rng(1)
path = synth.loop_path(nb + 1)
cfg = kc.LaserScanConfig(lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], "")
scans = [kc.LocalizedRangeScan(cfg, synth.cast_scan(world, p, P, rng), kc.Pose2(*p), kc.Pose2(*p), i, 0.0) for i, p in enumerate(path[:nb])]
true_q = path[nb - 1] + np.array([0.07, -0.04, 0.03])
q = kc.LocalizedRangeScan(cfg, synth.cast_scan(world, true_q, P, rng), kc.Pose2(*path[nb - 1]), kc.Pose2(*path[nb - 1]), nb, 0.0)
w = kc.Wrapper(kc.ScanMatcherConfig(), max_slots=4)
for _ in range(30):
    w.match_scan(q, scans, True, True)
def p50(f, n=300):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return np.percentile(np.array(ts) * 1e6, [50, 99])
print("Wrapper.match_scan p50/p99 us", p50(lambda: w.match_scan(q, scans, True, True)))
pool, starts, counts = pack_pool([q.point_readings()] + [s.point_readings() for s in scans])
poses = np.array([q.sensor_pose()])
m = w.matcher
args = (pool, starts, counts, np.array([0], np.int32), poses, np.array([0, nb], np.int32), np.arange(1, nb + 1, dtype=np.int32))
print("match_pool p50/p99 us", p50(lambda: m.match_pool(*args, True, True)))
print("match_pool coarse-only p50/p99 us", p50(lambda: m.match_pool(*args, True, False)))
if os.environ.get("YSM_TRACE"):
    m.match_pool(*args, True, True)
    print("==== through Wrapper.match_scan (content tags: scans come from the device-resident store)", file=sys.stderr, flush=True)
    for _ in range(3):
        w.match_scan(q, scans, True, True)
