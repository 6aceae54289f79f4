"""Developer quick check on a GPU box: CUDA path vs oracle on a few seeded batches."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from yag_slam_b200 import synth  # noqa: E402
from yag_slam_b200.matcher import ScanMatcherB200  # noqa: E402
import scenarios  # noqa: E402


def compare(name, cfg, batch, penalty, do_fine, max_slots=0):
    t0 = time.time()
    ref = scenarios.oracle_results(cfg, batch, penalty, do_fine)
    t_cpu = time.time() - t0
    m = ScanMatcherB200(cfg, max_slots=max_slots)
    m.match_pool(batch["pool"], batch["starts"], batch["counts"], batch["query_scan"], batch["query_pose"],
                 batch["base_ptr"], batch["base_idx"], penalty, do_fine)
    t0 = time.time()
    out = m.match_pool(batch["pool"], batch["starts"], batch["counts"], batch["query_scan"], batch["query_pose"],
                       batch["base_ptr"], batch["base_idx"], penalty, do_fine)
    t_gpu = time.time() - t0
    n = len(ref)
    resp_ok = (out["response"] == ref[:, 0])
    pose_ok = (out["x"] == ref[:, 1]) & (out["y"] == ref[:, 2]) & (out["heading"] == ref[:, 3])
    cov = out["cov"]
    rel = np.abs(cov - ref[:, 4:]) / np.maximum(np.abs(ref[:, 4:]), 1e-300)
    rel = np.where(ref[:, 4:] == cov, 0, rel)
    print(f"[{name}] n={n} response exact {resp_ok.sum()}/{n} pose exact {pose_ok.sum()}/{n} "
          f"cov max rel {rel.max():.3e} passes {np.bincount(out['n_passes'])} cpu {t_cpu:.3f}s gpu {t_gpu:.4f}s "
          f"launches {m.launch_count()}")
    bad = np.where(~(resp_ok & pose_ok))[0]
    for i in bad[:5]:
        print("   mismatch", i, "gpu", out["response"][i], out["x"][i], out["y"][i], out["heading"][i], out["n_ties"][i],
              "ref", ref[i, :4])
    m.close()
    return bool(resp_ok.all() and pose_ok.all() and rel.max() < 1e-5)


def main():
    w = synth.make_world()
    ok = True
    ok &= compare("cfg1-like P=360 1 base", None, scenarios.make_batch(w, 4, 360, 1, 1, perturb=(0.07, 0.03)), True, True)
    ok &= compare("seq P=720 10 base", None, scenarios.make_batch(w, 24, 720, 10, 2), True, True)
    ok &= compare("seq no-penalty coarse only", None, scenarios.make_batch(w, 8, 720, 10, 3), False, False)
    loop = dict(search_size=4.0, resolution=0.05)
    ok &= compare("loop P=720 10 base + degenerate", loop,
                  scenarios.make_batch(w, 40, 720, 10, 4, perturb=(1.0, 0.2), degenerate_frac=0.2), False, False)
    ok &= compare("loop shared query", loop,
                  scenarios.make_batch(w, 32, 720, 10, 5, perturb=(1.0, 0.2), shared_query=True), False, False)
    ok &= compare("multi-wave (slots=3)", None, scenarios.make_batch(w, 10, 360, 3, 6), True, True, max_slots=3)
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
