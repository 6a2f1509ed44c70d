#!/bin/bash
# parity of the two widenings (SUBGRID, MOVING_BOUNDARIES) on the GPU + the reference's lid-driven cavity in its own formulation
mkdir -p gpurun_out
exec > >(tee gpurun_out/next_rows.log) 2>&1
echo "=== parity"; timeout 900 python -m pytest tests -m gpu -q -x -k "subgrid or moving" 2>&1 | tail -3
for wl in d3q19_srt_fp32_256_cavity_mb d3q19_srt_fp32_256_cavity; do
  echo "=== bench $wl"; timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-e2e | tee gpurun_out/final_$wl.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'], d['roofline']['kernel'])"
done
echo "=== host cavity (MOVING_BOUNDARIES, D3Q19 SRT FP16S 128^3, 10000 steps)"; (cd fluidx3d_b200/host && timeout 120 bin/FluidX3D_CAVITY 2>&1 | tr '\r' '\n' | tail -3 | cut -c1-120)
