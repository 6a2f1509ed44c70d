#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session20.log) 2>&1
echo "=== cavity"; timeout 600 python bench.py --workload d3q19_srt_fp32_256_cavity | tee gpurun_out/bench_cavity.json | cut -c1-1200
echo "=== windtunnel full"; timeout 900 python bench.py --workload d3q27_trt_fp32_windtunnel_full --steps 50 --warmup 5 --no-cpu-baseline | tee gpurun_out/bench_windtunnel_full.json | cut -c1-1200
