#!/bin/bash
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --matches 20000 --no-latency --no-extras --no-cpu"
timeout 600 ncu --set full --clock-control none -k regex:k_tile_clear -s 30 -c 1 -f -o gpurun_out/r02zc_k_tile_clear $BENCH > gpurun_out/r02zc_k_tile_clear.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r02zc_k_tile_clear.ncu-rep --page raw --csv | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); hdr=r[0]; v=r[2]
for k in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','lts__t_sectors_op_write.sum','lts__t_sectors_srcunit_tex_op_write.sum']:
    if k in hdr: print(k, v[hdr.index(k)])
"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
timeout 600 python bench.py --steps 3 --warmup 2 --no-latency --no-extras --no-cpu > gpurun_out/r02zc_bench_$rep.json 2> gpurun_out/r02zc_bench_$rep.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02zc_bench_$rep.json').read().strip().splitlines()[-1])
print('rep=$rep value', round(d['value']), 'e2e', round(d['e2e']['value']))
PY
done
