#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session19.log) 2>&1
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "=== cavity"; timeout 600 python bench.py --workload d3q19_srt_fp32_256_cavity | tee gpurun_out/bench_cavity.json | cut -c1-900
echo "=== windtunnel full"; timeout 900 python bench.py --workload d3q27_trt_fp32_windtunnel_full --steps 50 --warmup 5 --no-cpu-baseline | tee gpurun_out/bench_windtunnel_full.json | cut -c1-900
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 | cut -c1-700
echo "=== host binaries"; (cd fluidx3d_b200/host && FX3D_BENCHMARK_SIZE=512 timeout 120 bin/FluidX3D 2>&1 | tail -4; timeout 300 bin/FluidX3D_POISEUILLE 2>&1 | tail -6)
