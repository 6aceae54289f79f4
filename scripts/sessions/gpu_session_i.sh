#!/bin/bash
# r01i (2 GPUs): parity with the new clear kernel, 1-GPU bench, 2-GPU torchrun bench (NCCL all-gather), reference arm under torchrun
TAG=${1:-r01i}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench n1 rc=$?"; cat gpurun_out/${TAG}_bench_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "bench n2 rc=$?"; cat gpurun_out/${TAG}_bench_n2.json; tail -5 gpurun_out/${TAG}_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_n2.json 2> gpurun_out/${TAG}_bench_ref_n2.err; echo "ref n2 rc=$?"; cat gpurun_out/${TAG}_bench_ref_n2.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
