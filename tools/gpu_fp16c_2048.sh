#!/bin/bash
# BASELINE configs[3]: D3Q19 FP16C 2048^3, 2x2x2 across 8 B200 (1024^3 + halo per GPU), against the same box's 1-GPU 1024^3 run
mkdir -p gpurun_out
exec > >(tee gpurun_out/fp16c_2048.log) 2>&1
echo "=== N=1 1024^3"
timeout 600 python bench.py --workload d3q19_srt_fp16c_1024 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e | tee gpurun_out/fp16c_1024_n1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['config']['global_grid'], d['value'], d['roofline']['frac'], d['ms_per_step'])"
echo "=== N=8 2048^3 split 2,2,2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --split 2,2,2 --workload d3q19_srt_fp16c_1024 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | grep '^{' | tee gpurun_out/fp16c_2048_n8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['config']['global_grid'], d['config']['domains'], d['value'], d['roofline']['frac'], d['ms_per_step'])"
echo "=== N=8 2048x2048x... split 1,2,4 (z/y stacking)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 8 --workload d3q19_srt_fp16c_1024 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | grep '^{' | tee gpurun_out/fp16c_1024x8_n8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['config']['global_grid'], d['config']['domains'], d['value'], d['roofline']['frac'], d['ms_per_step'])"
