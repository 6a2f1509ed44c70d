#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session22.log) 2>&1
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for wl in d3q19_srt_fp16s_512 d3q19_srt_fp16c_512 d3q19_srt_fp32_512 d3q27_trt_fp32_windtunnel d3q19_srt_fp32_256_cavity; do
  echo "=== bench $wl"; timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"
done
