"""Developer check of the resident latency kernel (k_match_resident) on a GPU box: single queries through
ysm_match_batch (n_matches = 1) against the oracle, handle interleaving, idle exit / relaunch, fallback
cases, the doorbell round-trip floor and the single-query p50."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from yag_slam_b200 import synth  # noqa: E402
from yag_slam_b200.distributed import slice_batch  # noqa: E402
from yag_slam_b200.matcher import ScanMatcherB200  # noqa: E402
import scenarios  # noqa: E402

LOOP = dict(search_size=4.0, resolution=0.05)


def singles(name, m, cfg, batch, penalty, do_fine, sleep=0.0):
    ref = scenarios.oracle_results(cfg, batch, penalty, do_fine)
    n = len(ref)
    bad = 0
    served0 = 0
    res = 0
    worst = 0.0
    for i in range(n):
        qs, qp, bp, bi = slice_batch(batch["query_scan"], batch["query_pose"], batch["base_ptr"], batch["base_idx"], i, i + 1)
        out = m.match_pool(batch["pool"], batch["starts"], batch["counts"], qs, qp, bp, bi, penalty, do_fine)
        res += m.last_work()["resident_requests"]
        r = ref[i]
        ok = (out["response"][0] == r[0] and out["x"][0] == r[1] and out["y"][0] == r[2] and out["heading"][0] == r[3])
        cov = out["cov"][0]
        scale = np.maximum(np.abs(r[4:]), np.sqrt(abs(r[4] * r[8])) * np.array([0, 1, 0, 1, 0, 0, 0, 0, 0]))
        rel = np.max(np.abs(cov - r[4:]) / np.maximum(scale, 1e-300))
        worst = max(worst, rel)
        if not ok or rel > 1e-5:
            bad += 1
            if bad <= 3:
                print("   mismatch", i, "gpu", out["response"][0], out["x"][0], out["y"][0], out["heading"][0],
                      out["n_passes"][0], out["n_ties"][0], "ref", r[:4], "cov rel", rel)
        if sleep:
            time.sleep(sleep)
    print(f"[{name}] n={n} bad={bad} resident-served={res} cov max rel {worst:.2e} launches {m.launch_count()}")
    return bad == 0


def main():
    w = synth.make_world()
    ok = True
    seq = ScanMatcherB200(None, max_slots=4, lanes=1)
    print("ping RTT us p50/p99 (first launches the kernel):", np.percentile(seq.ping(300)[5:], [50, 99]))
    ok &= singles("cfg1 P=360 nb=1 fine", seq, None, scenarios.make_batch(w, 40, 360, 1, 1, perturb=(0.07, 0.03)), True, True)
    ok &= singles("cfg2 P=720 nb=10 fine", seq, None, scenarios.make_batch(w, 30, 720, 10, 2), True, True)
    ok &= singles("coarse only no penalty", seq, None, scenarios.make_batch(w, 10, 720, 10, 3), False, False)
    ok &= singles("idle exit between calls", seq, None, scenarios.make_batch(w, 6, 360, 3, 7), True, True, sleep=0.01)
    loop = ScanMatcherB200(LOOP, max_slots=4, lanes=1)
    bl = scenarios.make_batch(w, 20, 720, 10, 4, perturb=(1.0, 0.2), degenerate_frac=0.2)
    ok &= singles("loop cfg + degenerate (fallback)", loop, LOOP, bl, False, False)
    # two handles interleaved (one resident kernel per device: they take turns)
    b1 = scenarios.make_batch(w, 8, 360, 2, 11)
    r1 = scenarios.oracle_results(None, b1, True, True)
    b2 = scenarios.make_batch(w, 8, 720, 5, 12, perturb=(1.0, 0.2))
    r2 = scenarios.oracle_results(LOOP, b2, False, False)
    bad = 0
    for i in range(8):
        for m, b, r, pen, fine in ((seq, b1, r1, True, True), (loop, b2, r2, False, False)):
            qs, qp, bp, bi = slice_batch(b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"], i, i + 1)
            o = m.match_pool(b["pool"], b["starts"], b["counts"], qs, qp, bp, bi, pen, fine)
            if not (o["response"][0] == r[i][0] and o["x"][0] == r[i][1] and o["y"][0] == r[i][2] and o["heading"][0] == r[i][3]):
                bad += 1
    print("[interleaved handles] bad =", bad)
    ok &= bad == 0
    # a batch through the general path while the resident kernel is alive, then singles again
    bb = scenarios.make_batch(w, 12, 360, 3, 13)
    rb = scenarios.oracle_results(None, bb, True, True)
    o = seq.match_pool(bb["pool"], bb["starts"], bb["counts"], bb["query_scan"], bb["query_pose"], bb["base_ptr"], bb["base_idx"], True, True)
    okb = bool((o["response"] == rb[:, 0]).all() and (o["x"] == rb[:, 1]).all() and (o["heading"] == rb[:, 3]).all())
    print("[batch after resident] ok =", okb)
    ok &= okb
    ok &= singles("singles after a batch", seq, None, scenarios.make_batch(w, 10, 360, 1, 14, perturb=(0.07, 0.03)), True, True)
    # latency
    from yag_slam_b200 import karto_compat as kc
    for P, nb in ((360, 1), (720, 10)):
        lp = synth.laser_params(P)
        rng = np.random.default_rng(1)
        path = synth.loop_path(nb + 1)
        cfg = kc.LaserScanConfig(lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], "")
        scans = [kc.LocalizedRangeScan(cfg, synth.cast_scan(w, p, P, rng), kc.Pose2(*p), kc.Pose2(*p), i, 0.0) for i, p in enumerate(path[:nb])]
        true_q = path[nb - 1] + np.array([0.07, -0.04, 0.03])
        q = kc.LocalizedRangeScan(cfg, synth.cast_scan(w, true_q, P, rng), kc.Pose2(*path[nb - 1]), kc.Pose2(*path[nb - 1]), nb, 0.0)
        wr = kc.Wrapper(kc.ScanMatcherConfig(), max_slots=4)
        for _ in range(50):
            wr.match_scan(q, scans, True, True)
        ts = []
        for _ in range(1000):
            t0 = time.perf_counter()
            wr.match_scan(q, scans, True, True)
            ts.append(time.perf_counter() - t0)
        print(f"Wrapper.match_scan P={P} nb={nb} p50/p99 us", np.percentile(np.array(ts) * 1e6, [50, 99]),
              "resident", wr.matcher.last_work()["resident_requests"])
        ts = []
        for _ in range(200):
            time.sleep(0.004)  # the kernel has left the device: cold call (launch inside)
            t0 = time.perf_counter()
            wr.match_scan(q, scans, True, True)
            ts.append(time.perf_counter() - t0)
        print(f"   cold (kernel relaunched per call) p50/p99 us", np.percentile(np.array(ts) * 1e6, [50, 99]))
        del wr
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
