#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session29.log) 2>&1
echo "=== subgrid parity"; timeout 900 python -m pytest tests -m gpu -q -x -k "subgrid" 2>&1 | tail -3
for wl in d3q19_srt_fp32_512_subgrid d3q19_srt_fp16s_512_subgrid; do
  echo "=== bench $wl"; timeout 600 python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline --no-e2e | tee gpurun_out/final_$wl.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'], d['roofline']['kernel'])"
done
