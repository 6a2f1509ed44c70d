#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session21.log) 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_pipe -s 4 -c 1 -o gpurun_out/prof21_windtunnel python bench.py --workload d3q27_trt_fp32_windtunnel --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu21.log 2>&1
tail -2 gpurun_out/ncu21.log | cut -c1-300
for v in 4 2; do echo "variant $v"; timeout 600 python bench.py --workload d3q27_trt_fp32_windtunnel --variant $v --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"; done
