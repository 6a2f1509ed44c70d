#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session14.log) 2>&1
for lib in _l1 _l2; do
  echo "=== parity lib=$lib"; FX3D_LIB=$PWD/fluidx3d_b200/libfx3d_cuda$lib.so timeout 600 python -m pytest tests -m gpu -q -x -k "(small and pipelined) or (medium and pipelined)" 2>&1 | tail -2
  for wl in d3q19_srt_fp32_512 d3q19_srt_fp16s_512 d3q19_srt_fp16c_512; do
    echo "=== bench $wl lib=$lib"; FX3D_LIB=$PWD/fluidx3d_b200/libfx3d_cuda$lib.so timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"
  done
done
echo "=== ncu fp32 l1"
FX3D_LIB=$PWD/fluidx3d_b200/libfx3d_cuda_l1.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_pipe -s 4 -c 1 -o gpurun_out/prof14_fp32_512_pipe_l1 python bench.py --workload d3q19_srt_fp32_512 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu14.log 2>&1
