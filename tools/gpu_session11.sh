#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session11.log) 2>&1
for lib in _s4 _t4 _s3 _g4; do
  echo "=== selftests lib=$lib"; FX3D_LIB=$PWD/fluidx3d_b200/libfx3d_cuda$lib.so timeout 600 python -m pytest tests -m gpu -q -x -k "packed or division or (small and pipelined) or (medium and pipelined)" 2>&1 | tail -3
  for wl in d3q19_srt_fp16s_512 d3q19_srt_fp16c_512 d3q19_srt_fp32_512; do
    echo "=== bench $wl lib=$lib variant=8"; FX3D_LIB=$PWD/fluidx3d_b200/libfx3d_cuda$lib.so timeout 600 python bench.py --workload $wl --variant 8 --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"
  done
done
echo "=== ncu s4 pipe"
FX3D_LIB=$PWD/fluidx3d_b200/libfx3d_cuda_s4.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_pipe -s 4 -c 1 -o gpurun_out/prof11_fp16s_512_pipe_s4 python bench.py --variant 8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu11.log 2>&1
