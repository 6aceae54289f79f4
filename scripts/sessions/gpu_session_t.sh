#!/bin/bash
# r01t (1 GPU): evidence at HEAD -- full GPU suite, smoke, bench both arms, launch list, ncu --set full of the
# grid-build kernels (exact candidate lists) and the pruned sweep
TAG=${1:-r01v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
timeout 500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
for k in k_tile_stamp k_find_valid k_sweep_pruned; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/${TAG}_$k -f python bench.py --steps 2 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
