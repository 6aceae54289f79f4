#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 correlative scan matcher.

Metric (BASELINE.json): scan matches/sec (+ p50 single-match latency) vs the Karto CPU matcher.
Workload at N=1 (BASELINE configs[1] shape): offline re-matching of a synthetic 2,000-scan
720-beam LiDAR log -- every scan is matched against its 10 running scans with yag_slam's
default Karto parameters (search 0.5 m, resolution 0.01 m, coarse 0.349/0.0349 rad, penalty on,
fine pass on). One "step" = one pass of the hot path over that batch of 2,000 independent
match queries. At N>1 every rank gets its own batch of the same shape (weak scaling), followed
by one NCCL all-gather of the 128-byte result records.

  python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...   # CPU reference arm (oracle port)

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "scan matches/sec"
UNIT = "matches/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--matches", type=int, default=2000)
    ap.add_argument("--beams", type=int, default=720)
    ap.add_argument("--base", type=int, default=10)
    ap.add_argument("--lanes", type=int, default=0,
                    help="matcher lanes (host threads + streams) per GPU; 0 = 3 when this rank has >= 4 host cores, else 2")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def pick_lanes(args):
    """Every lane is a host thread that plans passes and waits on its stream. Three overlap best, at
    N=1 (r01s: 307k -> 315k matches/s on 16 cores) and at N=8 with 4 cores per rank (r01u: 2.34M -> 2.39M);
    two when a rank has fewer than 4 host cores."""
    if args.lanes > 0:
        return args.lanes
    world = max(1, int(os.environ.get("WORLD_SIZE", "1")))
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    return 3 if cores // world >= 4 else 2


def workload_config(args):
    return {
        "workload": "cfg2-log-rematch: %d-scan %d-beam synthetic LiDAR log, each scan vs its %d running scans, "
                    "yag_slam default_config (search 0.5, res 0.01, coarse 0.349/0.0349, fine 0.00349), "
                    "penalty=True, do_fine=True" % (args.matches, args.beams, args.base),
        "matches_per_step_per_gpu": args.matches, "lanes": args.lanes,
        "beams": args.beams,
        "base_scans": args.base,
        "l2": "inputs larger than L2: each step touches ~%d correlation grids (16.6 MB slots, ~1 MB of lines each) "
              "+ the 46 MB point pool, >> 126 MB L2" % args.matches,
    }


def make_workload(args, rank):
    from yag_slam_b200 import synth
    world = synth.make_world()
    return synth.make_match_batch(world, args.matches, args.beams, args.base, seed=2 + 1000 * rank, perturb=(0.1, 0.05))


class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[0]))
                smax.append(float(p[1]))
            except Exception:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:  # region shorter than one sample: take the nearest sample
            for ts, line in self.rows[-3:]:
                p = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(p[0]))
                    smax.append(float(p[1]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one sweep launch, from the committed ncu --set full
    capture of this same command (profiles/sweep_traffic.json); None when no capture is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "sweep_traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_read_per_launch"] + t["dram_bytes_write_per_launch"]), t["source"]
    except Exception:
        return None, None


def host_cores():
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would starve the CPU arm: the thread count is taken from the affinity mask instead)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(cfg, b, n_sample, n_threads):
    """Oracle timed on the first n_sample matches of the workload (checker used as the CPU arm)."""
    from oracle import oracle
    from yag_slam_b200.distributed import slice_batch
    qs, qp, bp, bi = slice_batch(b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"], 0, n_sample)
    t0 = time.perf_counter()
    out = oracle.match_batch(cfg, b["pool"], b["starts"], b["counts"], qs, qp, bp, bi, True, True, n_threads)
    return time.perf_counter() - t0, out


def run_reference(args, rank, world):
    """CPU reference arm: the reference's own implementation of the path is the external wheel
    karto_scanmatcher==1.0.0 (not installable here, source absent), so this times the oracle port
    of it on the host cores, all threads, on a bounded sample of the same workload per step."""
    if rank != 0:
        return
    b = make_workload(args, 0)
    cores = host_cores()
    # calibrate the per-step sample to ~3 s of wall clock
    dt, _ = cpu_sample(None, b, min(args.matches, 2 * cores), cores)
    rate = (2 * cores) / max(dt, 1e-6)
    n_sample = int(max(cores, min(args.matches, rate * 3.0)))
    for _ in range(args.warmup):
        cpu_sample(None, b, n_sample, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample(None, b, n_sample, cores)
    dt = time.perf_counter() - t0
    value = n_sample * args.steps / dt
    sample = "first %d of the %d matches of the workload per step, %d OpenMP threads, one matcher per thread" % (
        n_sample, args.matches, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 sums / f64 poses",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "reference wheel karto_scanmatcher 1.0.0 unavailable; baseline is the in-repo "
                                 "restatement (oracle/karto_oracle.c), incl. Karto's full-grid memset per match"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from yag_slam_b200 import _capi, distributed
    from yag_slam_b200.matcher import ScanMatcherB200

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    b = make_workload(args, rank)
    n = args.matches
    m = ScanMatcherB200(None, device=local_rank, lanes=args.lanes)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    dpool = torch.from_numpy(b["pool"]).to(dev)
    hpool = torch.from_numpy(b["pool"]).pin_memory()
    res = np.zeros(n, dtype=_capi.RESULT_DTYPE)
    acc = {"sweep_ms": 0.0, "build_ms": 0.0, "reduce_ms": 0.0, "total_ms": 0.0, "lookups": 0, "launches": 0,
           "offset_entries": 0, "poses": 0, "h2d": 0, "d2h": 0, "issued": 0, "pruned_launches": 0, "count": False,
           "base_points": 0}

    def step(pool):
        out = m.match_pool(pool, b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"],
                           b["base_idx"], True, True, stream=sp, out=res)
        if acc["count"]:
            k, w = m.last_kernel_ms(), m.last_work()
            acc["sweep_ms"] += k["sweep"]; acc["build_ms"] += k["build"]; acc["reduce_ms"] += k["reduce"]
            acc["total_ms"] += k["total"]
            acc["lookups"] += w["lattice_lookups"]; acc["launches"] += w["sweep_launches"]
            acc["offset_entries"] += w["offset_entries"]; acc["poses"] += w["poses"]
            acc["h2d"] += w["h2d_bytes"]; acc["d2h"] += w["d2h_bytes"]
            acc["issued"] += w["lookups_issued"]; acc["pruned_launches"] += w["pruned_sweep_launches"]
            acc["base_points"] += w["base_points"]
        if world > 1:
            # one all-gather of best poses / responses over NVLink; weak scaling: every
            # rank contributes its own n records
            t = torch.from_numpy(out.view(np.float64).reshape(n, 16)).to(dev)
            g = torch.empty((world * n, 16), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(g, t)
            return g
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(pool, K):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(K):
            step(pool)
        e1.record(stream)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        # the matcher's lanes run on their own streams: the host clock brackets them all (the call
        # is synchronous), the events see the caller's stream
        ms = max(e0.elapsed_time(e1), (t1 - t0) * 1e3)
        wall_ms = (t1 - t0) * 1e3
        v = torch.tensor([ms, wall_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v[0]), float(v[1]), t0, t1

    # ---- device-resident arm ("value") -----------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step(dpool)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = m.launch_count()
    ms, wall_ms, t0, t1 = timed(dpool, args.steps)
    launches = m.launch_count() - launches0
    clocks = sampler.stop(t0, t1) if sampler else None
    value = world * n * args.steps / (ms * 1e-3)

    # ---- roofline pass: the same steps with per-kernel CUDA events on the launching stream. Kernel
    # timing keeps the whole batch on one lane, so the dominant kernel is timed running alone. ------
    m.set_debug(_capi.DEBUG_TIME_KERNELS)
    step(dpool)
    acc["count"] = True
    for _ in range(args.steps):
        step(dpool)
    acc["count"] = False
    m.set_debug(0)
    snap = dict(acc)

    # ---- end-to-end arm: pinned host pool, H2D + D2H inside the timed region ------------------------
    for _ in range(max(args.warmup, 3)):
        step(hpool)
    ems, ewall_ms, _, _ = timed(hpool, args.steps)
    w = m.last_work()
    acc["h2d"], acc["d2h"] = w["h2d_bytes"] * args.steps, w["d2h_bytes"] * args.steps
    e2e_value = world * n * args.steps / (max(ems, ewall_ms) * 1e-3)

    # ---- p50 single-match latency through the public API ----------------------------------------
    lat = {}
    if rank == 0 and not args.no_latency:
        lat = latency_probe(local_rank, with_cpu=(world == 1 and not args.no_cpu))

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        dt, ref = cpu_sample(None, b, min(n, 8 * cores), cores)
        n_sample = int(max(cores, min(n, (8 * cores / max(dt, 1e-6)) * 12.0)))
        dt, ref = cpu_sample(None, b, n_sample, cores)
        # the timed GPU results must equal the CPU results on the sample (bit-exact pose/response)
        exact = bool((res["response"][:n_sample] == ref[:, 0]).all() and (res["x"][:n_sample] == ref[:, 1]).all()
                     and (res["y"][:n_sample] == ref[:, 2]).all() and (res["heading"][:n_sample] == ref[:, 3]).all())
        t1c, _ = cpu_sample(None, b, min(n, 32), 1)
        cpu = {"value": n_sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d of the %d matches of the timed workload, %d OpenMP threads, one matcher per thread "
                         "(%.1f s of CPU wall clock)" % (n_sample, n, cores, dt),
               "single_thread_ms_per_match": 1e3 * t1c / min(n, 32),
               "gpu_results_bit_exact_on_sample": exact,
               "note": "reference wheel karto_scanmatcher 1.0.0 unavailable; baseline is the in-repo restatement"}

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    traffic, traffic_src = measured_traffic()
    sweep_bytes = snap["lookups"] * 1 + snap["offset_entries"] * 4 + snap["poses"] * 4
    sweep_s = snap["sweep_ms"] * 1e-3
    achieved = sweep_bytes / sweep_s / 1e9 if sweep_s > 0 else None
    # the grid build (FindValidPoints + SmearPoint): SURVEY 8d counts 2 B per byte-max op, S = P_base * K^2
    ksz = int(m.dims()["kernel_size"])
    stamp_bytes = snap["base_points"] * ksz * ksz * 2
    build_s = snap["build_ms"] * 1e-3
    stamp_achieved = stamp_bytes / build_s / 1e9 if build_s > 0 else None
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8 sums / f64 poses", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": acc["h2d"] // max(args.steps, 1),
                "d2h_bytes_per_step": acc["d2h"] // max(args.steps, 1) + n * 128,
                "ms_per_step": max(ems, ewall_ms) / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "kernel": ("k_sweep_pruned" if snap["pruned_launches"] else "k_sweep_lattice") +
                      " (CorrelateScan/GetResponse coarse sweep)",
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": sweep_bytes / max(snap["launches"], 1),
            "avg_launch_ms": snap["sweep_ms"] / max(snap["launches"], 1), "launches": snap["launches"],
            "lookups_per_s": snap["lookups"] / sweep_s if sweep_s > 0 else None,
            # exact zero-row pruning: the algorithmic lookups above are Karto's; this many were really issued
            "lookups_issued_frac": (snap["issued"] / snap["lookups"]) if snap["pruned_launches"] and snap["lookups"] else 1.0,
            "timed_in": "separate pass of the same %d steps on one lane with per-kernel CUDA events" % args.steps,
            "share_of_step": snap["sweep_ms"] / max(snap["total_ms"], 1e-9),
            "build_share": snap["build_ms"] / max(snap["total_ms"], 1e-9),
            "reduce_share": snap["reduce_ms"] / max(snap["total_ms"], 1e-9),
        },
        "roofline_build": {
            "kernel": "k_find_valid + k_tile_stamp (AddScans/FindValidPoints/SmearPoint)", "bound": "hbm",
            "achieved": stamp_achieved, "peak": peak, "unit": "GB/s",
            "frac": (stamp_achieved / peak) if stamp_achieved else None,
            "algorithmic_bytes": "2 B x base points x K^2 (K = %d), base points before FindValidPoints" % ksz,
            "avg_launch_ms": snap["build_ms"] / max(snap["launches"], 1),
        },
        "cpu_baseline": cpu,
    }
    out.update(lat)
    print(json.dumps(out))


def latency_probe(device, with_cpu=False):
    """p50 of single match_scan calls through the reference-facing API (Wrapper.match_scan):
    cfg 1 (360 beams, 1 base scan) and cfg 2 shape (720 beams, 10 running scans). with_cpu: the
    same single queries on the CPU oracle, one thread (the cpu_baseline leg)."""
    from yag_slam_b200 import karto_compat as kc
    from yag_slam_b200 import synth
    world = synth.make_world()
    w = kc.Wrapper(kc.ScanMatcherConfig(), device=device, max_slots=4)
    out = {}
    for name, P, nb, reps in (("cfg1", 360, 1, 1000), ("cfg2", 720, 10, 500)):
        lp = synth.laser_params(P)
        rng = np.random.default_rng(1)
        path = synth.loop_path(nb + 1)
        cfg = kc.LaserScanConfig(lp[0], lp[1], lp[2], lp[3], lp[4], lp[5], "")
        scans = [kc.LocalizedRangeScan(cfg, synth.cast_scan(world, p, P, rng), kc.Pose2(*p), kc.Pose2(*p), i, 0.0)
                 for i, p in enumerate(path[:nb])]
        true_q = path[nb - 1] + np.array([0.07, -0.04, 0.03])
        q = kc.LocalizedRangeScan(cfg, synth.cast_scan(world, true_q, P, rng), kc.Pose2(*path[nb - 1]),
                                  kc.Pose2(*path[nb - 1]), nb, 0.0)
        q.point_readings()
        for s in scans:
            s.point_readings()
        for _ in range(50):
            w.match_scan(q, scans, True, True)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            w.match_scan(q, scans, True, True)
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e6
        out["p50_latency_us_" + name] = float(np.percentile(ts, 50))
        out["p99_latency_us_" + name] = float(np.percentile(ts, 99))
        if with_cpu:
            from oracle.oracle import KartoOracle
            o = KartoOracle(None)
            qp, bp = q.point_readings(), [s.point_readings() for s in scans]
            pose = q.sensor_pose()
            cts = []
            for _ in range(12):
                t0 = time.perf_counter()
                ref = o.match(qp, pose, bp, True, True)
                cts.append(time.perf_counter() - t0)
            r = w.match_scan(q, scans, True, True)
            exact = (r.response == ref[0] and (r.best_pose.x, r.best_pose.y, r.best_pose.yaw) == ref[1])
            out["cpu_p50_latency_us_" + name] = float(np.percentile(np.array(cts[2:]) * 1e6, 50))
            out["latency_speedup_" + name] = out["cpu_p50_latency_us_" + name] / out["p50_latency_us_" + name]
            out["latency_result_bit_exact_" + name] = bool(exact)
    return out


def main():
    args = parse_args()
    args.lanes = pick_lanes(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback for the product path "
                           "(use --impl reference for the CPU arm)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
