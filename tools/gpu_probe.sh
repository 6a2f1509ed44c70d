#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/probe.log) 2>&1
nproc; free -g | head -2
for split in 1,1,2 2,1,1; do
echo "=== probe split $split"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tools/dist_probe.py $split 2>&1 | grep "^rank"
done
