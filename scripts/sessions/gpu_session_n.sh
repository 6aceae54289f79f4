#!/bin/bash
# r01n (8 GPUs): weak-scaling bench at N=8 (NCCL all-gather of the records), host core count
TAG=${1:-r01u}
mkdir -p gpurun_out
{ nproc; nvidia-smi -L | wc -l; free -g | head -2; } > gpurun_out/${TAG}_host.txt 2>&1; cat gpurun_out/${TAG}_host.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_n8.json; tail -5 gpurun_out/${TAG}_bench_n8.err
