#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session8.log) 2>&1
echo "=== pytest gpu (subset)"; time timeout 900 python -m pytest tests -m gpu -q -x -k "small or packed or division" 2>&1 | tail -4
for lib in "" _f3 _f4 _f5; do
  for v in 8 4; do
    for wl in d3q19_srt_fp16s_512 d3q19_srt_fp32_512 d3q19_srt_fp16c_512; do
      echo "=== bench $wl lib=$lib variant=$v"; FX3D_LIB=$PWD/fluidx3d_b200/libfx3d_cuda$lib.so timeout 600 python bench.py --workload $wl --variant $v --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"
    done
  done
done
echo "=== ncu f4 pipe"
FX3D_LIB=$PWD/fluidx3d_b200/libfx3d_cuda_f4.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_pipe -s 4 -c 1 -o gpurun_out/prof8_fp16s_512_pipe_f4 python bench.py --variant 8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu8.log 2>&1
