#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session18.log) 2>&1
echo "=== parity subset"; timeout 900 python -m pytest tests -m gpu -q -x -k "small or medium or golden or decomposed" 2>&1 | tail -3
for wl in d3q19_srt_fp16s_512 d3q19_srt_fp16c_512 d3q19_srt_fp32_512 d3q19_srt_fp32_256 d3q27_trt_fp32_windtunnel; do
  echo "=== bench $wl"; timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"
done
