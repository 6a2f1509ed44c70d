#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session26.log) 2>&1
echo "=== parity q27"; timeout 900 python -m pytest tests -m gpu -q -x -k "q27" 2>&1 | tail -3
for wl in d3q27_trt_fp32_windtunnel d3q27_trt_fp32_windtunnel_full; do
  echo "=== bench $wl"; timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --no-e2e | tee gpurun_out/final_$wl.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"
done
