"""Developer probe: real (warm-L2, in-stream) per-kernel times of one bench-shaped call, first wave,
single lane (YSM_TRACE_GPU=1 makes the library print CUDA-event deltas)."""
import os, sys
os.environ["YSM_TRACE_GPU"] = "1"  # read once by the library, at its first call
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from yag_slam_b200 import synth
from yag_slam_b200.matcher import ScanMatcherB200
world = synth.make_world()
b = synth.make_match_batch(world, 2000, 720, 10, seed=2, perturb=(0.1, 0.05))
m = ScanMatcherB200(None, lanes=1)
if os.environ.get("YSM_TRACE_DEBUG"):
    m.set_debug(int(os.environ["YSM_TRACE_DEBUG"]))  # e.g. 64 = YSM_DEBUG_NO_HALF_LISTS
dpool = torch.from_numpy(b["pool"]).cuda()
args = (dpool, b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"], True, True)
for _ in range(3):
    m.match_pool(*args)
sys.stderr.flush()
print("==== last call", file=sys.stderr, flush=True)
m.match_pool(*args)
