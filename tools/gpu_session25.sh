#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session25.log) 2>&1
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for wl in d3q19_srt_fp16s_512 d3q19_srt_fp16c_512 d3q19_srt_fp32_512 d3q19_srt_fp32_256 d3q19_srt_fp16s_256 d3q19_srt_fp32_256_cavity d3q27_trt_fp32_windtunnel; do
  echo "=== bench $wl"; timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-e2e | tee gpurun_out/final_$wl.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"
done
echo "=== default bench"; timeout 900 python bench.py | tee gpurun_out/final_default.json | cut -c1-400
echo "=== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_r01_final.csv | cut -c1-300
echo "=== ncu fp32"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_tma -s 4 -c 1 -o gpurun_out/prof25_fp32_512_tma python bench.py --workload d3q19_srt_fp32_512 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu25.log 2>&1
