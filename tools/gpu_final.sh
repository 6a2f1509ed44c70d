#!/bin/bash
# what the driver runs at round end, in one go: GPU test suite, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
exec > >(tee gpurun_out/final.log) 2>&1
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (default)"; timeout 900 python bench.py | tee gpurun_out/bench_default.json
echo "=== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 | cut -c1-300
echo "=== x-halo fp32 auto"; timeout 300 python tools/xhalo_probe.py 0 fp32
echo "=== host binary"; (cd fluidx3d_b200/host && FX3D_BENCHMARK_SIZE=512 timeout 120 bin/FluidX3D 2>&1 | tail -2)
