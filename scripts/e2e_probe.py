"""Developer probe: match_pool step time with a device-resident pool vs a pinned host pool
(wave-sliced upload on / off via YSM_NO_SLICED_UPLOAD)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from yag_slam_b200 import synth
from yag_slam_b200.matcher import ScanMatcherB200
world = synth.make_world()
b = synth.make_match_batch(world, 2000, 720, 10, seed=2, perturb=(0.1, 0.05))
m = ScanMatcherB200(None, lanes=2)
dpool = torch.from_numpy(b["pool"]).cuda()
hpool = torch.from_numpy(b["pool"]).pin_memory()
res = np.zeros(2000, dtype=m.match_pool.__globals__["_capi"].RESULT_DTYPE)
def run(pool, k=20):
    for _ in range(3):
        m.match_pool(pool, b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"], True, True, out=res)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k):
        m.match_pool(pool, b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"], True, True, out=res)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
for rep, kb in enumerate(("1024", "2048", "4096", "8192")):
    os.environ.pop("YSM_NO_SLICED_UPLOAD", None)
    os.environ["YSM_UPLOAD_PIECE_KB"] = kb
    a = run(dpool); c = run(hpool)
    os.environ["YSM_NO_SLICED_UPLOAD"] = "1"
    d = run(hpool)
    print("piece KB", kb, "ms/step: device pool %.3f | pinned sliced %.3f | pinned whole-pool %.3f | h2d MB %.1f" % (a, c, d, m.last_work()["h2d_bytes"] / 1e6))
