"""Developer probe (GPU box): throughput of the batched chain finder at the cfg-3 shape -- a 2,000-scan
pose graph (BASELINE cfg 2 trajectory), 4,096 query scans -- against the CPU oracle restatement of
the reference's Python (oracle/chains_oracle.py) on a bounded sample; checks equality on the sample."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from oracle import chains_oracle as co
from yag_slam_b200 import chains, synth

n, nq = 2000, 4096
rng = np.random.default_rng(3)
path = synth.loop_path(n, step=0.25)[:, :2] + rng.normal(0, 0.05, (n, 2))
seq = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1)
loops = np.array([[i - 283, i] for i in range(300, n, 40)])
ptr, idx = chains.adjacency_csr(n, np.concatenate([seq, loops]))
queries = rng.integers(0, n, nq).astype(np.int32)
for _ in range(3):
    cs = chains.find_chains_batch(path, ptr, idx, queries, 3, 10)
ts, kms = [], []
for _ in range(10):
    t0 = time.perf_counter()
    cs = chains.find_chains_batch(path, ptr, idx, queries, 3, 10)
    ts.append(time.perf_counter() - t0)
    kms.append(cs.kernel_ms)
ns = 64
t0 = time.perf_counter()
ref = co.find_chains_batch(path, path, ptr, idx, queries[:ns], 3, 10)
tcpu = time.perf_counter() - t0
ok = (cs.query_chain_ptr[:ns + 1] == ref[0]).all() and (cs.members[:len(ref[2])] == ref[2]).all()
print(json.dumps({"what": "find_possible_loop_closure_chains, batched", "vertices": n, "queries": nq,
                  "chains": int(cs.n_chains), "members": int(len(cs.members)),
                  "gpu_call_ms_p50": float(np.median(ts) * 1e3), "gpu_kernel_ms_p50": float(np.median(kms)),
                  "queries_per_s_e2e": nq / float(np.median(ts)),
                  "cpu_port_queries_per_s": ns / tcpu, "cpu_sample": "%d queries, pure-Python restatement, 1 core" % ns,
                  "equal_on_sample": bool(ok)}))
