#!/bin/bash
# r01z2 (N GPUs): cfg 5 strong scaling -- 100k relocalisation queries sharded over N ranks + all-gather, sharded ray-walk
N=${1:-4}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/reloc_dist_probe.py > gpurun_out/r01z_reloc_n$N.json 2> gpurun_out/r01z_reloc_n$N.err; echo "rc=$?"
cat gpurun_out/r01z_reloc_n$N.json; tail -5 gpurun_out/r01z_reloc_n$N.err
