"""Generates tests/golden/graph_checkpoint.bin: a GraphSlam checkpoint written by the REFERENCE's own
GraphSlam.to_file (yag_slam/graph_slam.py:91-100, unmodified), from a 12-scan sequential mapping run
over the synthetic world (matcher compute = the CPU oracle behind the karto_scanmatcher-compatible
value types; the wheel is absent). The reference cannot travel to the GPU box, so the file is
committed.  Re-run:  python tests/golden/make_graph_checkpoint.py"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "shims"))
sys.path.insert(0, "/root/reference")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")

import numpy as np  # noqa: E402

from yag_slam_b200 import karto_compat, synth  # noqa: E402
from test_host_cpu import _OracleWrapper  # noqa: E402

mod = types.ModuleType("karto_scanmatcher")
for n in ("Pose2", "LaserScanConfig", "LocalizedRangeScan", "ScanMatcherConfig", "create_occupancy_grid"):
    setattr(mod, n, getattr(karto_compat, n))
mod.Wrapper = _OracleWrapper
sys.modules["karto_scanmatcher"] = mod

import yag_slam.graph_slam as gs  # noqa: E402
import yag_slam.models as models  # noqa: E402
import yag_slam.scan_matching as sm  # noqa: E402

N, P, L = 12, 360, 5
world = synth.make_world()
lp = synth.laser_params(P)
rng = np.random.default_rng(33)
path = synth.loop_path(N, step=0.25)
odom = synth.noisy_odometry(path, rng, 0.01, 0.005)
slam = gs.GraphSlam(sm.Scan2DMatcherCpp({}), None, scan_buffer_len=L)
for k in range(N):
    scan = models.LocalizedRangeScan(np.round(synth.cast_scan(world, path[k], P, rng), 4), lp[0], lp[1], lp[2], lp[3],
                                     lp[4], lp[5], *odom[k])
    slam.process_scan(scan)
slam.to_file(os.path.join(HERE, "graph_checkpoint.bin"))
print("wrote", os.path.getsize(os.path.join(HERE, "graph_checkpoint.bin")), "bytes")
