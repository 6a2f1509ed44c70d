#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session23.log) 2>&1
echo "=== parity"; timeout 900 python -m pytest tests -m gpu -q -x -k "small or medium or golden or decomposed" 2>&1 | tail -4
for wl in d3q19_srt_fp16s_512 d3q19_srt_fp32_512 d3q19_srt_fp16c_512 d3q27_trt_fp32_windtunnel d3q19_srt_fp32_256; do
  for v in 8 16; do
  echo "=== bench $wl variant $v"; timeout 600 python bench.py --workload $wl --variant $v --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"
  done
done
echo "=== ncu tma fp16s"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_tma -s 4 -c 1 -o gpurun_out/prof23_fp16s_512_tma python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu23.log 2>&1
tail -2 gpurun_out/ncu23.log | cut -c1-200
