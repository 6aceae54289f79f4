#!/bin/bash
# r02x (1 GPU): host-side phase trace of the throughput path (one lane, 333-match waves like the bench's 3-lane split)
mkdir -p gpurun_out
YSM_TRACE=1 timeout 300 python - > gpurun_out/r02x_host_trace.txt 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch, time
from yag_slam_b200 import synth
from yag_slam_b200.matcher import ScanMatcherB200
world = synth.make_world()
b = synth.make_match_batch(world, 2000, 720, 10, seed=2, perturb=(0.1, 0.05))
m = ScanMatcherB200(None, lanes=1, max_slots=333)
dpool = torch.from_numpy(b["pool"]).cuda()
args = (dpool, b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"], True, True)
for _ in range(2):
    m.match_pool(*args)
sys.stderr.flush()
print("==== last call", file=sys.stderr, flush=True)
t0=time.perf_counter(); m.match_pool(*args); t1=time.perf_counter()
print("==== 2000 matches in %.2f ms" % ((t1-t0)*1e3), file=sys.stderr)
PY
sed -n '/==== last call/,$p' gpurun_out/r02x_host_trace.txt | head -90
