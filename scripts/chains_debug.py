import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from test_chains_cpu import load_case
from yag_slam_b200 import chains
d = load_case("default")
cs = chains.find_chains_batch(d["pose_xy"], d["adj_ptr"], d["adj_idx"], d["queries"], d["dist"], d["min_chain"], hash_xy=d["hash_xy"])
print("qcp equal", (cs.query_chain_ptr == d["query_chain_ptr"]).all(), "n_chains", cs.n_chains, len(d["chain_ptr"]) - 1, "members", len(cs.members), len(d["members"]))
shown = 0
for i, q in enumerate(d["queries"]):
    ref = [d["members"][d["chain_ptr"][c]:d["chain_ptr"][c + 1]].tolist() for c in range(d["query_chain_ptr"][i], d["query_chain_ptr"][i + 1])]
    got = cs.chains_of(i) if i + 1 < len(cs.query_chain_ptr) else None
    if got != ref:
        print("query", i, "vertex", q, "\n  ref", ref, "\n  got", got)
        shown += 1
        if shown >= 4:
            break
