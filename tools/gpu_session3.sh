#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session3.log) 2>&1
echo "=== pytest gpu"; time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
for lib in "" _mb2 _mb4; do
  for wl in d3q19_srt_fp16s_512 d3q19_srt_fp32_512 d3q19_srt_fp16c_512; do
    echo "=== bench $wl lib=$lib"; FX3D_LIB=$PWD/fluidx3d_b200/libfx3d_cuda$lib.so timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"
  done
done
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_v4 -s 6 -c 1 -o gpurun_out/prof3_fp16s_512 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_v4 -s 6 -c 1 -o gpurun_out/prof3_fp32_512 python bench.py --workload d3q19_srt_fp32_512 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu3b.log 2>&1
