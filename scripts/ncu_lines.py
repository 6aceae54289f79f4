"""Per-CUDA-source-line instruction / stall totals from an ncu report (needs -lineinfo + --import-source on).
usage: python scripts/ncu_lines.py <report.ncu-rep> [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None
agg = collections.OrderedDict()
cur = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    d = dict(zip(hdr, r))
    # rows with a line number start a source line; SASS rows below it carry the metrics
    if r[0].strip():
        cur = (int(r[0]), r[1].strip())
        agg.setdefault(cur, [0, 0, 0])
    try:
        inst = int(r[hdr.index("Instructions Executed")] or 0)
        samp = int(r[hdr.index("# Samples")] or 0)
    except ValueError:
        continue
    if cur is not None and r[2].strip():
        agg[cur][0] += inst
        agg[cur][1] += samp
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
for (ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5d  inst %5.1f%%  samples %5.1f%%  %s" % (ln, 100.0 * v[0] / tot_i, 100.0 * v[1] / tot_s, src[:110]))
