#!/bin/bash
# r02zzc (1 GPU): final HEAD of round 2: both bench arms, every section
mkdir -p gpurun_out
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02zzc_bench_reference.json 2> gpurun_out/r02zzc_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/r02zzc_bench.json 2> gpurun_out/r02zzc_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zzc_bench.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r02zzc_bench_reference.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ref', round(r['value']), 'ratio', round(d['e2e']['value']/r['value'],1))
print('roofline', {k:round(d['roofline'][k],3) for k in ('frac','lsu_frac','share_of_timed_kernels','avg_launch_ms')}, 'build', round(d['roofline_build']['frac'],3), 'step', round(d['roofline_step']['frac'],3))
print({k:round(v,1) for k,v in d.items() if 'p50_latency' in k or 'speedup' in k})
print('cfg3', round(d['cfg3']['matches_per_s']), 'cfg4 p50', round(d['cfg4']['p50_latency_us']), 'seq', round(d['cfg2_sequential']['front_end']['scans_per_s']), 'rays', round(d['cfg5_final_map']['rays_per_s']/1e6), 'M/s')
print([k for k in d if k.endswith('_error')])
PY
