#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/probe.log) 2>&1
for split in 1,1,2 2,1,1; do
echo "=== probe split $split"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tools/dist_probe.py $split 2>&1 | grep "^rank 0"
done
echo "=== dist_check"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py 2>&1 | grep "DIST_CHECK\|MISMATCH"
echo "=== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | grep '^{' | tee gpurun_out/scale_2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['roofline']['frac'], d['ms_per_step'], d['e2e'])"
