#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session6.log) 2>&1
echo "=== pytest gpu (fast subset)"; time timeout 900 python -m pytest tests -m gpu -q -x -k "small or packed or division or decomposed" 2>&1 | tail -5
for lib in "" _a _b; do
  for v in 4 2; do
    for wl in d3q19_srt_fp16s_512 d3q19_srt_fp32_512 d3q19_srt_fp16c_512; do
      echo "=== bench $wl lib=$lib K=$v"; FX3D_LIB=$PWD/fluidx3d_b200/libfx3d_cuda$lib.so timeout 600 python bench.py --workload $wl --variant $v --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"
    done
  done
done
