#!/bin/bash
# what the driver runs at round end, in one go: GPU test suite, smoke, the default bench line, the reference arm -- plus a few extras
mkdir -p gpurun_out
exec > >(tee gpurun_out/final.log) 2>&1
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (default)"; timeout 900 python bench.py | tee gpurun_out/bench_default.json | cut -c1-1500
echo "=== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 | cut -c1-200
for wl in d3q19_srt_fp32_512_subgrid d3q19_srt_fp16s_512_subgrid d3q19_srt_fp16c_1024; do
  echo "=== bench $wl"; timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --no-e2e | tee gpurun_out/final_$wl.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'], d['roofline']['kernel'])"
done
echo "=== host binary with SUBGRID (wind tunnel scene, D3Q19 SRT FP16S)"; (cd fluidx3d_b200/host && timeout 40 bin/FluidX3D_WINDTUNNEL 2>&1 | tr '\r' '\n' | tail -4 | cut -c1-120)
