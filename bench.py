#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 correlative scan matcher.

Metric (BASELINE.json): scan matches/sec (+ p50 single-match latency) vs the Karto CPU matcher.

Workload of `value` / `e2e` (BASELINE configs[4], the same per-match shape as configs[1]): batched
relocalisation of 100,000 independent (query, 10 running scans) pairs sampled from a synthetic 2,000-scan
720-beam LiDAR log, yag_slam's default Karto parameters (search 0.5 m, resolution 0.01 m, coarse 0.349/0.0349
rad, penalty on, fine pass on). One "step" = one pass of the hot path over those 100,000 queries. Under
--gpus N the SAME 100,000 queries are sharded contiguously over the ranks (strong scaling), followed by one
NCCL all-gather of the 128-byte result records.

On rank 0 at N = 1 the line also carries the other BASELINE configs: the 2,000-match log re-matching batch
(`cfg2_rematch`, round 1's headline), cfg 3 (4,096 loop-closure chains per query), cfg 4 (high-resolution single
query), cfg 1 / cfg 2 single-query p50 through Wrapper.match_scan, and cfg 2 as defined (the reference's
unmodified GraphSlam.process_scan driving the CUDA Wrapper).

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
LSU_BOUND = 9.0e12  # ~148 SMs x 1 wavefront/clk x ~1.9 GHz x <= 32 one-byte lookups per wavefront
LOOP = dict(search_size=4.0, resolution=0.05)
CFG4 = dict(resolution=0.005, search_size=1.0, fine_search_angle_resolution=0.00175, smear_deviation=0.05)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--matches", type=int, default=100000)
    ap.add_argument("--beams", type=int, default=720)
    ap.add_argument("--base", type=int, default=10)
    ap.add_argument("--lanes", type=int, default=0,
                    help="matcher lanes (host threads + streams) per GPU; 0 = 3 when this rank has >= 4 host cores, else 2")
    ap.add_argument("--max-slots", type=int, default=0,
                    help="correlation grids resident per rank (all lanes together; 0 = the library's HBM budget): "
                         "the wave size of the throughput path")
    ap.add_argument("--grid-gb", type=float, default=0.0,
                    help="HBM budget of the correlation-grid slots per rank in GiB (0 = the library default, 16)")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline workload (no cfg 2/3/4 sections)")
    return ap.parse_args()


def pick_lanes(args):
    """Every lane is a host thread that plans passes and waits on its stream. Three overlap best, at
    N=1 (r01s: 307k -> 315k matches/s on 16 cores) and at N=8 with 4 cores per rank (r01u: 2.34M -> 2.39M);
    two when a rank has fewer than 4 host cores."""
    if args.lanes > 0:
        return args.lanes
    world = max(1, int(os.environ.get("WORLD_SIZE", "1")))
    return 3 if host_cores() // world >= 4 else 2


def workload_config(args):
    return {
        "workload": "cfg5-relocalisation: %d independent (query, %d running scans) pairs sampled from a synthetic "
                    "2000-scan %d-beam LiDAR log, pose perturbation U(+-0.2 m, +-0.15 rad), yag_slam default_config "
                    "(search 0.5, res 0.01, coarse 0.349/0.0349, fine 0.00349), penalty=True, do_fine=True; the same "
                    "queries sharded contiguously over the GPUs (strong scaling)" % (args.matches, args.base, args.beams),
        "matches_per_step": args.matches, "lanes": args.lanes, "beams": args.beams, "base_scans": args.base,
        "l2": "inputs larger than L2: each step touches %d correlation grids (16.6 MB slots, ~1 MB of lines each) "
              "+ a %.1f GB point pool, >> 126 MB L2" % (args.matches, args.matches * args.beams * 16 / 1e9),
    }


def make_workload(args, n_use=None, oracle_readings=False):
    """The seeded cfg-5 batch (identical on every rank). oracle_readings: point readings by the oracle -- the
    CPU reference arm must not load the product library."""
    from yag_slam_b200 import synth
    readings = None
    if oracle_readings:
        from oracle import oracle
        readings = oracle.point_readings
    world = synth.make_world()
    n_log = max(2000, args.base + 1)
    return synth.make_relocalisation_batch(world, args.matches, args.beams, args.base, seed=5, n_log=n_log, n_use=n_use,
                                           readings=readings)


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


def cpu_sample(cfg, b, n_sample, n_threads, penalty=True, do_fine=True):
    """Oracle timed on the first n_sample matches of a workload (checker used as the CPU arm)."""
    from oracle import oracle
    from yag_slam_b200.distributed import slice_batch
    qs, qp, bp, bi = slice_batch(b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"], 0, n_sample)
    t0 = time.perf_counter()
    out = oracle.match_batch(cfg, b["pool"], b["starts"], b["counts"], qs, qp, bp, bi, penalty, do_fine, n_threads)
    return time.perf_counter() - t0, out


def run_reference(args, rank, world):
    """CPU reference arm: the reference's own implementation of the path is the external wheel
    karto_scanmatcher==1.0.0 (not installable here, source absent), so this times the oracle port
    of it on the host cores, all threads, on a bounded sample of the same workload per step. The workload's
    point readings come from the oracle too: this arm never loads the product library."""
    if rank != 0:
        return
    cores = host_cores()
    # a prefix of the workload large enough for ~3 s of CPU time per step at any plausible core count
    n_avail = min(args.matches, max(64 * cores, 256))
    b = make_workload(args, n_use=n_avail, oracle_readings=True)
    dt, _ = cpu_sample(None, b, min(n_avail, 2 * cores), cores)
    rate = min(n_avail, 2 * cores) / max(dt, 1e-6)
    n_sample = int(max(min(cores, n_avail), min(n_avail, rate * 3.0)))
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
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8 sums / f64 poses",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "reference wheel karto_scanmatcher 1.0.0 unavailable; baseline is the in-repo "
                                 "restatement (oracle/karto_oracle.c), incl. Karto's full-grid memset per match"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from yag_slam_b200 import _capi
    from yag_slam_b200.distributed import shard_range, slice_batch
    from yag_slam_b200.matcher import ScanMatcherB200

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    t_gen = time.perf_counter()
    b = make_workload(args)
    t_gen = time.perf_counter() - t_gen
    n = args.matches
    lo, hi = shard_range(n, rank, world)  # strong scaling: this rank's contiguous share of the same batch
    qs, qp, bp, bi = slice_batch(b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"], lo, hi)
    nloc = hi - lo
    m = ScanMatcherB200(None, device=local_rank, lanes=args.lanes, max_slots=args.max_slots,
                        max_grid_bytes=int(args.grid_gb * (1 << 30)))
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    dpool = torch.from_numpy(b["pool"]).to(dev)
    hpool = torch.from_numpy(b["pool"]).pin_memory()
    res = np.zeros(nloc, dtype=_capi.RESULT_DTYPE)
    per = -(-n // world)
    send = torch.zeros((per, 16), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * per, 16), dtype=torch.float64, device=dev) if world > 1 else None
    acc = {"sweep_ms": 0.0, "build_ms": 0.0, "reduce_ms": 0.0, "total_ms": 0.0, "lookups": 0, "launches": 0,
           "offset_entries": 0, "poses": 0, "h2d": 0, "d2h": 0, "issued": 0, "pruned_launches": 0, "count": False,
           "base_points": 0, "valid_points": 0, "fine_lookups": 0}

    def step(pool):
        out = m.match_pool(pool, b["starts"], b["counts"], qs, qp, bp, bi, True, True, stream=sp, out=res)
        if acc["count"]:
            k, w = m.last_kernel_ms(), m.last_work()
            acc["sweep_ms"] += k["sweep"]; acc["build_ms"] += k["build"]; acc["reduce_ms"] += k["reduce"]
            acc["total_ms"] += k["total"]
            acc["lookups"] += w["lattice_lookups"]; acc["launches"] += w["sweep_launches"]
            acc["offset_entries"] += w["offset_entries"]; acc["poses"] += w["poses"]
            acc["h2d"] += w["h2d_bytes"]; acc["d2h"] += w["d2h_bytes"]
            acc["issued"] += w["lookups_issued"]; acc["pruned_launches"] += w["pruned_sweep_launches"]
            acc["base_points"] += w["base_points"]; acc["valid_points"] += w["valid_base_points"]
            acc["fine_lookups"] += w["fine_lookups"]
        if world > 1:
            # one all-gather of best poses / responses over NVLink: every rank ends up with all n records
            send[:nloc].copy_(torch.from_numpy(out.view(np.float64).reshape(nloc, 16)), non_blocking=False)
            dist.all_gather_into_tensor(gathered, send)
            return gathered
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
    value = n * args.steps / (ms * 1e-3)
    first = res[:min(nloc, 4096)].copy()

    # ---- roofline pass: the same work with per-kernel CUDA events on the launching stream. Kernel
    # timing keeps the whole batch on one lane, so the dominant kernel is timed running alone. ------
    m.set_debug(_capi.DEBUG_TIME_KERNELS)
    step(dpool)
    acc["count"] = True
    for _ in range(min(args.steps, 2)):
        step(dpool)
    acc["count"] = False
    m.set_debug(0)
    snap = dict(acc)

    # ---- end-to-end arm: pinned host pool, H2D + D2H inside the timed region ------------------------
    for _ in range(max(args.warmup, 3)):
        step(hpool)
    ems, ewall_ms, _, _ = timed(hpool, args.steps)
    w = m.last_work()
    e2e_h2d, e2e_d2h = w["h2d_bytes"], w["d2h_bytes"] + nloc * 128
    e2e_value = n * args.steps / (max(ems, ewall_ms) * 1e-3)
    e2e_same = bool(res[:len(first)].tobytes() == first.tobytes())
    ksz = int(m.dims()["kernel_size"])
    m.close()
    del dpool, hpool

    extras = {}
    cpu = None
    if rank == 0 and world == 1:
        cores = host_cores()
        if not args.no_cpu:
            # ---- CPU baseline of the headline workload (bounded sample, all host threads) -------------
            dt, ref = cpu_sample(None, b, min(n, 8 * cores), cores)
            n_sample = int(max(cores, min(n, 4096, (8 * cores / max(dt, 1e-6)) * 12.0)))
            dt, ref = cpu_sample(None, b, n_sample, cores)
            exact = bool((first["response"][:n_sample] == ref[:, 0]).all() and (first["x"][:n_sample] == ref[:, 1]).all()
                         and (first["y"][:n_sample] == ref[:, 2]).all() and (first["heading"][:n_sample] == ref[:, 3]).all())
            t1c, _ = cpu_sample(None, b, min(n, 32), 1)
            cpu = {"value": n_sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "first %d of the %d matches of the timed workload, %d OpenMP threads, one matcher per thread "
                             "(%.1f s of CPU wall clock)" % (n_sample, n, cores, dt),
                   "single_thread_ms_per_match": 1e3 * t1c / min(n, 32),
                   "gpu_results_bit_exact_on_sample": exact,
                   "note": "reference wheel karto_scanmatcher 1.0.0 unavailable; baseline is the in-repo restatement"}
        del b
        if not args.no_extras:
            extras.update(section(lambda: bench_rematch(args, local_rank), "cfg2_rematch"))
            extras.update(section(lambda: bench_cfg3(local_rank, cores, not args.no_cpu), "cfg3"))
            extras.update(section(lambda: bench_cfg4(local_rank, not args.no_cpu), "cfg4"))
            extras.update(section(lambda: bench_sequential(local_rank), "cfg2_sequential"))
            extras.update(section(lambda: bench_final_map(local_rank, not args.no_cpu), "cfg5_final_map"))
            extras.update(section(lambda: bench_chain_finder(local_rank, not args.no_cpu), "f2_chain_finder"))
            extras.update(section(lambda: bench_map_match(local_rank, not args.no_cpu), "f3_map_match"))
        if not args.no_latency:
            extras.update(section(lambda: latency_probe(local_rank, with_cpu=not args.no_cpu), "latency"))

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    traffic, traffic_src = measured_traffic()
    sweep_bytes = snap["lookups"] * 1 + snap["offset_entries"] * 4 + snap["poses"] * 4
    sweep_s = snap["sweep_ms"] * 1e-3
    achieved = sweep_bytes / sweep_s / 1e9 if sweep_s > 0 else None
    # the grid build (FindValidPoints + SmearPoint): SURVEY 8d counts 2 B per byte-max op, S = P_valid * K^2 with
    # P_valid = base points that survive FindValidPoints and the ROI test
    stamp_bytes = snap["valid_points"] * ksz * ksz * 2
    build_s = snap["build_ms"] * 1e-3
    stamp_achieved = stamp_bytes / build_s / 1e9 if build_s > 0 else None
    # whole step: B = L + 8 T + 2 S + 8 R + 128 per match, over the timed passes
    n_timed = min(args.steps, 2) * nloc
    step_bytes = (snap["lookups"] + snap["fine_lookups"] + 8 * snap["offset_entries"] + stamp_bytes
                  + 8 * n_timed * (args.base + 1) * args.beams + 128 * n_timed)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8 sums / f64 poses", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_h2d), "d2h_bytes_per_step": int(e2e_d2h),
                "ms_per_step": max(ems, ewall_ms) / args.steps, "same_records_as_resident_arm": e2e_same,
                "note": "per rank: pinned host pool -> only the scans the rank's waves reference are uploaded"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "timed_region_s": ms * 1e-3,
        "workload_generate_s": t_gen,
        "roofline": {
            "kernel": ("k_sweep_pruned" if snap["pruned_launches"] else "k_sweep_lattice") +
                      " (CorrelateScan/GetResponse coarse sweep)",
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": sweep_bytes / max(snap["launches"], 1),
            "avg_launch_ms": snap["sweep_ms"] / max(snap["launches"], 1), "launches": snap["launches"],
            "lookups_per_s": snap["lookups"] / sweep_s if sweep_s > 0 else None,
            # the physically binding limit: L1/LSU wavefronts, <= 32 one-byte lookups each
            "lsu_bound_lookups_per_s": LSU_BOUND,
            "lsu_frac": (snap["lookups"] / sweep_s / LSU_BOUND) if sweep_s > 0 else None,
            # exact zero-row pruning: the algorithmic lookups above are Karto's; this many were really issued
            "lookups_issued_frac": (snap["issued"] / snap["lookups"]) if snap["pruned_launches"] and snap["lookups"] else 1.0,
            "timed_in": "separate pass of %d steps on one lane with per-kernel CUDA events" % min(args.steps, 2),
            # share_of_step is against the single-lane wall time of the timed pass, host gaps included (the GPU idles
            # ~ a quarter of it there; three lanes fill those gaps in the headline run). The figure to hold against the
            # ncu launch list (profiles/*_launches.csv: sweep / (sweep + find_valid + stamp + reduce)) is this one:
            "share_of_timed_kernels": snap["sweep_ms"] / max(snap["sweep_ms"] + snap["build_ms"] + snap["reduce_ms"], 1e-9),
            "share_of_step": snap["sweep_ms"] / max(snap["total_ms"], 1e-9),
            "build_share": snap["build_ms"] / max(snap["total_ms"], 1e-9),
            "reduce_share": snap["reduce_ms"] / max(snap["total_ms"], 1e-9),
        },
        "roofline_build": {
            "kernel": "k_find_valid + k_tile_stamp_lists (AddScans/FindValidPoints/SmearPoint)", "bound": "hbm",
            "achieved": stamp_achieved, "peak": peak, "unit": "GB/s",
            "frac": (stamp_achieved / peak) if stamp_achieved else None,
            "algorithmic_bytes": "2 B x P_valid x K^2 (K = %d); P_valid = %d of %d base points survive FindValidPoints "
                                 "and the ROI test" % (ksz, snap["valid_points"], snap["base_points"]),
            "avg_launch_ms": snap["build_ms"] / max(snap["launches"], 1),
        },
        "roofline_step": {
            "what": "whole step, SURVEY 8d bytes per match B = L + 8T + 2S + 8R + 128 over the single-lane timed passes",
            "bound": "hbm", "achieved": step_bytes / (snap["total_ms"] * 1e-3) / 1e9 if snap["total_ms"] > 0 else None,
            "peak": peak, "unit": "GB/s",
            "frac": step_bytes / (snap["total_ms"] * 1e-3) / 1e9 / peak if snap["total_ms"] > 0 else None,
            "bytes_per_match": step_bytes / max(n_timed, 1),
        },
        "cpu_baseline": cpu,
    }
    out.update(extras)
    print(json.dumps(out))


def section(fn, name):
    """An extra section must never take the headline line down with it."""
    try:
        return fn()
    except Exception as e:  # noqa: BLE001
        return {name + "_error": "%s: %s" % (type(e).__name__, str(e)[:300])}


def _timed_batches(m, b, penalty, do_fine, reps):
    import torch
    ts = []
    out = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = m.match_pool(b["pool"], b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"],
                           b["base_idx"], penalty, do_fine)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), out


def _kernel_roofline(m, b, penalty, do_fine):
    """One pass with per-kernel CUDA events: lookups/s and algorithmic GB/s of the coarse sweep."""
    from yag_slam_b200 import _capi
    m.set_debug(_capi.DEBUG_TIME_KERNELS)
    m.match_pool(b["pool"], b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"],
                 penalty, do_fine)
    k, w = m.last_kernel_ms(), m.last_work()
    m.set_debug(0)
    peak, _ = measured_peak()
    by = w["lattice_lookups"] + 4 * w["offset_entries"] + 4 * w["poses"]
    s = k["sweep"] * 1e-3
    return {"kernel": "k_sweep_pruned" if w["pruned_sweep_launches"] else "k_sweep_lattice", "bound": "hbm",
            "achieved": by / s / 1e9 if s > 0 else None, "peak": peak, "unit": "GB/s",
            "frac": by / s / 1e9 / peak if s > 0 else None, "lookups": w["lattice_lookups"],
            "lookups_per_s": w["lattice_lookups"] / s if s > 0 else None,
            "lsu_frac": w["lattice_lookups"] / s / LSU_BOUND if s > 0 else None,
            "sweep_ms": k["sweep"], "build_ms": k["build"], "reduce_ms": k["reduce"], "total_ms": k["total"],
            "sweep_launches": w["sweep_launches"]}


def bench_rematch(args, device):
    """Round 1's headline workload, kept for continuity: the 2,000-scan log re-matched in one batch."""
    import torch
    from yag_slam_b200 import synth
    from yag_slam_b200.matcher import ScanMatcherB200
    world = synth.make_world()
    b = synth.make_match_batch(world, 2000, args.beams, args.base, seed=2, perturb=(0.1, 0.05))
    m = ScanMatcherB200(None, device=device, lanes=args.lanes)
    dpool = torch.from_numpy(b["pool"]).to("cuda:%d" % device)
    bd = dict(b, pool=dpool)
    _timed_batches(m, bd, True, True, 3)
    t_dev, _ = _timed_batches(m, bd, True, True, 20)
    t_host, _ = _timed_batches(m, b, True, True, 10)
    m.close()
    return {"cfg2_rematch": {"workload": "2000-scan log, each scan vs its 10 running scans, one batch (round 1's headline)",
                             "matches_per_s": 2000 / t_dev, "ms_per_batch": 1e3 * t_dev,
                             "e2e_matches_per_s": 2000 / t_host, "reps": 20}}


def bench_cfg3(device, cores, with_cpu):
    """BASELINE cfg 3 (graph_slam.py:217-236 as one batch): ONE query against 4,096 candidate chains of 10
    scans, loop config (search 4.0, res 0.05), response expansion on, 10% degenerate chains; coarse pass
    (penalty off, no fine) for every chain, then the sequential-config refine (fine on) of the survivors."""
    from yag_slam_b200 import synth
    from yag_slam_b200.matcher import ScanMatcherB200
    world = synth.make_world()
    n = 4096
    b = synth.make_match_batch(world, n, 720, 10, 3, perturb=(1.0, 0.2), degenerate_frac=0.1, shared_query=True)
    m = ScanMatcherB200(LOOP, device=device)
    _timed_batches(m, b, False, False, 2)
    t, out = _timed_batches(m, b, False, False, 7)
    roof = _kernel_roofline(m, b, False, False)
    m.close()
    keep = np.where(out["response"] >= 0.35)[0]  # GraphSlam.min_response_coarse
    res = {"workload": "1 query (720 beams) x 4096 chains x 10 scans, loop config, expansion on, 10% degenerate; "
                       "coarse (penalty off, no fine)",
           "matches_per_s": n / t, "ms_per_query": 1e3 * t, "passes_histogram": np.bincount(out["n_passes"]).tolist(),
           "survivors": int(len(keep)), "roofline": roof}
    if len(keep):
        # refine the survivors with the sequential matcher at the coarse pose (graph_slam.py:233-236)
        nb = np.diff(b["base_ptr"])
        bp = np.concatenate([[0], np.cumsum(nb[keep])]).astype(np.int32)
        bi = np.concatenate([b["base_idx"][b["base_ptr"][i]:b["base_ptr"][i + 1]] for i in keep]).astype(np.int32)
        # tmpscan.corrected_pose = coarse pose (graph_slam.py:233-234): the query's readings move with it. For this
        # TIMING section they are re-posed by a rigid transform of the original readings (Karto re-evaluates
        # r cos / r sin at the new pose; no parity is claimed for this stage here, tests/ cover the refine pass).
        q0 = int(b["query_scan"][0])
        qp = b["pool"][b["starts"][q0]:b["starts"][q0] + b["counts"][q0]]
        p0 = b["query_pose"][0]
        c0, s0 = np.cos(-p0[2]), np.sin(-p0[2])
        loc = (qp - p0[:2]) @ np.array([[c0, s0], [-s0, c0]])
        poses = np.stack([out["x"][keep], out["y"][keep], out["heading"][keep]], axis=1)
        extra = []
        for px, py, ph in poses:
            c, s_ = np.cos(ph), np.sin(ph)
            extra.append(loc @ np.array([[c, s_], [-s_, c]]) + np.array([px, py]))
        n0, cnt = len(b["pool"]), len(qp)
        pool2 = np.concatenate([b["pool"]] + extra)
        starts2 = np.concatenate([b["starts"], n0 + cnt * np.arange(len(keep))]).astype(np.int32)
        counts2 = np.concatenate([b["counts"], np.full(len(keep), cnt)]).astype(np.int32)
        qs2 = (len(b["starts"]) + np.arange(len(keep))).astype(np.int32)
        b2 = dict(pool=pool2, starts=starts2, counts=counts2, query_scan=qs2, query_pose=np.ascontiguousarray(poses),
                  base_ptr=bp, base_idx=bi)
        ms = ScanMatcherB200(None, device=device)
        _timed_batches(ms, b2, False, True, 1)
        t2, _ = _timed_batches(ms, b2, False, True, 3)
        ms.close()
        res["refine_matches_per_s"] = len(keep) / t2
        res["ms_per_query_with_refine"] = 1e3 * (t + t2)
    if with_cpu:
        ns = min(n, 2 * cores)
        dt, ref = cpu_sample(LOOP, b, ns, cores, False, False)
        res["cpu_matches_per_s"] = ns / dt
        res["cpu_cores"] = cores
        res["bit_exact_on_cpu_sample"] = bool((out["response"][:ns] == ref[:, 0]).all() and (out["x"][:ns] == ref[:, 1]).all()
                                              and (out["heading"][:ns] == ref[:, 3]).all())
    return {"cfg3": res}


def bench_cfg4(device, with_cpu):
    """BASELINE cfg 4: 1,081 beams, resolution 0.005, search 1.0, fine 0.00175, smear 0.05 (= 10 x res: Karto's
    order-dependent smear), 1 base scan: 68 MB grid, 101 x 101 x 21 poses, 231.6 M lookups per match."""
    from yag_slam_b200 import synth
    from yag_slam_b200.matcher import ScanMatcherB200
    world = synth.make_world()
    b = synth.make_match_batch(world, 1, 1081, 1, 4, perturb=(0.1, 0.03))
    m = ScanMatcherB200(CFG4, device=device, max_slots=2, lanes=1)
    args = (b["pool"], b["starts"], b["counts"], b["query_scan"], b["query_pose"], b["base_ptr"], b["base_idx"], True, True)
    for _ in range(5):
        m.match_pool(*args)
    ts = []
    for _ in range(60):
        t0 = time.perf_counter()
        out = m.match_pool(*args)
        ts.append(time.perf_counter() - t0)
    roof = _kernel_roofline(m, b, True, True)
    d = m.dims()
    m.close()
    res = {"workload": "1 query, 1081 beams, res 0.005, search 1.0, fine 0.00175, smear 0.05, 1 base scan",
           "grid_bytes": int(d["grid_bytes"]), "p50_latency_us": float(np.percentile(np.array(ts) * 1e6, 50)),
           "p99_latency_us": float(np.percentile(np.array(ts) * 1e6, 99)), "roofline": roof}
    if with_cpu:
        from oracle import oracle
        cts = []
        for _ in range(3):
            t0 = time.perf_counter()
            ref = oracle.match_batch(CFG4, *args[:7], True, True, 1)
            cts.append(time.perf_counter() - t0)
        res["cpu_p50_latency_us"] = float(np.median(cts) * 1e6)
        res["latency_speedup"] = res["cpu_p50_latency_us"] / res["p50_latency_us"]
        res["bit_exact"] = bool(out["response"][0] == ref[0, 0] and out["x"][0] == ref[0, 1] and out["y"][0] == ref[0, 2]
                                and out["heading"][0] == ref[0, 3])
    return {"cfg4": res}


def bench_final_map(device, with_cpu):
    """The second half of BASELINE cfg 5: occupancy grid of the final map (karto create_occupancy_grid over the
    2,000-scan 720-beam log, reference graph_slam.py:341-342) and its ray-walk (reference raytracing.py:63-92,
    one 1,439-angle sweep from each of 1,024 free cells, the splicing caller's shape). Wall clock through the
    public API: host scan arrays in, host rays out (29 MB D2H inside the timed call)."""
    from yag_slam_b200 import occupancy, raytracing, synth
    world = synth.make_world()
    log = synth.make_scan_log(world, 2000, 720, seed=2)
    args = (log["poses"], log["lasers"], log["ranges"], log["beam_ptr"], 0.05, 12.0)
    occupancy.occupancy_grid_from_arrays(*args, device=device).close()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        g = occupancy.occupancy_grid_from_arrays(*args, device=device)
        ts.append(time.perf_counter() - t0)
        if len(ts) < 3:
            g.close()
    img = np.array(g.image)
    ang = np.arange(1439) * (360.0 / 1439) - 180.0
    free = np.argwhere(img == 255)
    starts = free[np.random.default_rng(0).choice(len(free), 1024, replace=False)][:, ::-1].astype(np.float64)
    raytracing.raytrace_many(g, ang, starts)
    rts = []
    for _ in range(5):
        t0 = time.perf_counter()
        rays = raytracing.raytrace_many(g, ang, starts)
        rts.append(time.perf_counter() - t0)
    dt = float(np.median(rts))
    nrays = rays.shape[0] * rays.shape[1]
    cells = float(rays[..., 4].astype(np.float64).sum())  # one map byte per visited cell (unit steps)
    res = {"workload": "final map of the 2000-scan 720-beam log at 0.05 m; 1024 starts x 1439 angles ray-walk",
           "map_wh": [int(g.width), int(g.height)], "occupancy_grid_ms": float(np.median(ts) * 1e3),
           "beams_per_s": float(len(log["ranges"]) / np.median(ts)), "raywalk_ms": dt * 1e3, "rays_per_s": nrays / dt,
           "mean_ray_cells": cells / nrays,
           "roofline": {"kernel": "k_raywalk (trace_ray)", "bound": "hbm",
                        "achieved": (cells + 20.0 * nrays) / dt / 1e9, "unit": "GB/s",
                        "algorithmic_bytes": "1 B per visited cell + 20 B per ray written; the call is bound by the "
                                             "dependent-load chain of each ray and the D2H of the rays, not by HBM",
                        "timed": "wall clock of the API call, D2H of the rays included"}}
    peak, _src = measured_peak()
    res["roofline"]["peak"] = peak
    res["roofline"]["frac"] = res["roofline"]["achieved"] / peak
    g.close()
    if with_cpu:
        from oracle import oracle
        ns = 16  # bounded sample: 16 sweeps of 1,439 rays, single thread (the reference's numba walk is not parallel)
        t0 = time.perf_counter()
        ref = oracle.raywalk_sweep_many(img, ang, starts[:ns])
        cdt = time.perf_counter() - t0
        res["cpu_rays_per_s"] = ns * len(ang) / cdt
        res["cpu_cores"] = 1
        res["raywalk_speedup"] = res["rays_per_s"] / res["cpu_rays_per_s"]
        res["rays_bit_exact_on_sample"] = bool((rays[:ns].view(np.uint32) == ref.view(np.uint32)).all())
        t0 = time.perf_counter()
        og = oracle.occupancy_grid(*args)
        res["cpu_occupancy_grid_ms"] = (time.perf_counter() - t0) * 1e3
        res["occupancy_image_equal"] = bool(og is not None and og["image"].shape == img.shape and (og["image"] == img).all())
    return {"cfg5_final_map": res}


def bench_chain_finder(device, with_cpu):
    """find_possible_loop_closure_chains (reference graph_slam.py:274-304) for 4,096 query
    vertices of a 2,000-vertex pose graph in one call of k_chain_find; CPU = the restatement of the reference's
    Python on a bounded sample, one core."""
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
    res = {"workload": "2000-vertex pose graph, 4096 query vertices, one call", "chains": int(cs.n_chains),
           "members": int(len(cs.members)), "call_ms_p50": float(np.median(ts) * 1e3),
           "kernel_ms_p50": float(np.median(kms)), "queries_per_s_e2e": nq / float(np.median(ts))}
    if with_cpu:
        from oracle import chains_oracle as co
        ns = 64
        t0 = time.perf_counter()
        ref = co.find_chains_batch(path, path, ptr, idx, queries[:ns], 3, 10)
        res["cpu_queries_per_s"] = ns / (time.perf_counter() - t0)
        res["cpu_sample"] = "%d queries, restatement of the reference's Python, 1 core" % ns
        res["equal_on_sample"] = bool((cs.query_chain_ptr[:ns + 1] == ref[0]).all() and
                                      (cs.members[:len(ref[2])] == ref[2]).all())
    return {"f2_chain_finder": res}


def bench_map_match(device, with_cpu):
    """4,096 720-beam queries against ONE resident correlation grid made from a map image
    (loop-matcher configuration, coarse only); CPU = the oracle on a bounded sample, one core."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_map_cpu import map_queries  # workload generator shared with the parity tests
    from yag_slam_b200 import synth
    from yag_slam_b200.matcher import DEFAULTS_LOOP, MapMatcherB200
    world = synth.make_world()
    n = 4096
    img, off, pool, starts, counts, qs, guess, truth = map_queries(world, n, 720, 9, 0.05, perturb=(1.0, 0.1))
    m = MapMatcherB200(DEFAULTS_LOOP, img, off, 0, device)
    for _ in range(3):
        out = m.match_map(pool, starts, counts, qs, guess, False, False)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        out = m.match_map(pool, starts, counts, qs, guess, False, False)
        ts.append(time.perf_counter() - t0)
    err = np.hypot(out["x"] - truth[:, 0], out["y"] - truth[:, 1])
    res = {"workload": "4096 queries x 720 beams vs a resident map grid, loop-matcher config, coarse only",
           "map_shape": list(img.shape), "call_ms_p50": float(np.median(ts) * 1e3),
           "matches_per_s_e2e": n / float(np.median(ts)), "median_position_error_m": float(np.median(err))}
    m.close()
    if with_cpu:
        from oracle import oracle
        ns = 48
        o = oracle.KartoMapOracle(DEFAULTS_LOOP, img, off, 0)
        t0 = time.perf_counter()
        ref = o.match_many(pool, starts, counts, qs[:ns], guess[:ns], False, False)
        res["cpu_matches_per_s"] = ns / (time.perf_counter() - t0)
        res["cpu_sample"] = "%d queries, 1 core" % ns
        res["bit_exact_on_sample"] = bool(all((out[k][:ns] == ref[:, c]).all()
                                              for k, c in (("response", 0), ("x", 1), ("y", 2), ("heading", 3))))
    return {"f3_map_match": res}


def bench_sequential(device):
    """BASELINE cfg 2 as SURVEY 8d defines it: the reference's UNMODIFIED GraphSlam.process_scan
    (yag_slam/graph_slam.py:306-339, from baseline/_ref) driving the CUDA Wrapper over a 2,000-scan 720-beam
    trajectory, scan_buffer_len = 10: front end only (loop_matcher = None) and with the loop matcher."""
    from harness import refslam
    from yag_slam_b200 import synth
    if refslam.reference_path() is None:
        return {"cfg2_sequential": {"unavailable": "reference package not installed under baseline/_ref"}}
    world = synth.make_world()
    n, beams = 2000, 720
    traj = refslam.make_trajectory(world, n, beams, seed=2)
    out = {}
    for name, with_loop in (("front_end", False), ("with_loop_matcher", True)):
        r = refslam.run_sequential(refslam.import_reference(), world, n, beams, with_loop=with_loop, traj=traj)
        err = np.hypot(r["poses"][:, 0] - r["truth"][:, 0], r["poses"][:, 1] - r["truth"][:, 1])
        out[name] = {"scans_per_s": n / r["total_s"], "total_s": r["total_s"],
                     "match_p50_us": float(np.percentile(r["match_s"] * 1e6, 50)),
                     "match_p99_us": float(np.percentile(r["match_s"] * 1e6, 99)),
                     "match_share_of_loop": float(r["match_s"].sum() / r["total_s"]),
                     "loops_closed": int(r["closed"]), "median_position_error_m": float(np.median(err))}
    out["workload"] = ("reference GraphSlam.process_scan, unmodified, on karto_compat.Wrapper; 2000 scans x 720 beams, "
                       "scan_buffer_len 10; the rest of the loop is the reference's Python (tiny_tf / sba_cpp stand-ins)")
    return {"cfg2_sequential": out}


def latency_probe(device, with_cpu=False):
    """p50 of single match_scan calls through the reference-facing API (Wrapper.match_scan):
    cfg 1 (360 beams, 1 base scan) and cfg 2 shape (720 beams, 10 running scans): back to back (the resident
    kernel stays on the device), and 'cold' (a pause longer than its idle time before every call, so each call
    relaunches it). with_cpu: the same single queries on the CPU oracle, one thread (the cpu_baseline leg)."""
    from yag_slam_b200 import karto_compat as kc
    from yag_slam_b200 import synth
    world = synth.make_world()
    w = kc.Wrapper(kc.ScanMatcherConfig(), device=device, max_slots=4)
    out = {}
    try:
        rtt = w.matcher.ping(300)
        out["doorbell_round_trip_us_p50"] = float(np.percentile(rtt[10:], 50))
    except Exception as e:  # noqa: BLE001
        out["doorbell_round_trip_error"] = str(e)[:200]
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
        out["latency_path_" + name] = "resident kernel" if w.matcher.last_work()["resident_requests"] else "launch per call"
        cold = []
        for _ in range(100):
            time.sleep(0.004)
            t0 = time.perf_counter()
            w.match_scan(q, scans, True, True)
            cold.append(time.perf_counter() - t0)
        out["p50_latency_us_%s_cold" % name] = float(np.percentile(np.array(cold) * 1e6, 50))
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
