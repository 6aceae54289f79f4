"""Summarise ncu artefacts brought back in gpurun_out/ into markdown under profiles/.
usage: python scripts/summarize_profiles.py <tag> [kernel-regex for the --set full report]"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def launches(tag):
    path = os.path.join(G, tag + "_launches.csv")
    if not os.path.exists(path):
        return ""
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d["Metric Name"] != "gpu__time_duration.sum":
                continue
            k = d["Kernel Name"].split("(")[0]
            v = float(d["Metric Value"].replace(",", ""))
            u = d["Metric Unit"]
            v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    out = ["| kernel | launches | total us | avg us | share % |", "|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("| %s | %d | %.1f | %.1f | %.1f |" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
    return "\n".join(out)


def full(tag, name):
    rep = os.path.join(G, "%s_%s.ncu-rep" % (tag, name))
    if not os.path.exists(rep):
        return ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = ["| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |",
           "|---|---|" + "---|" * len(data)]
    name_i = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out.append("| %s | %s | %s |" % (k, units[i], " | ".join(r[i] for r in data)))
    kn = data[0][name_i].split("(")[0] if name_i is not None else name
    return "Kernel: `%s`\n\n" % kn + "\n".join(out)


if __name__ == "__main__":
    tag = sys.argv[1]
    names = sys.argv[2:]
    print("## %s launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)\n" % tag)
    print(launches(tag))
    for n in names:
        print("\n## %s `ncu --set full` of %s\n" % (tag, n))
        print(full(tag, n))
