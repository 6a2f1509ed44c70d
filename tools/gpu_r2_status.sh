#!/bin/bash
# round-2 status run: a short bench first (a hang ends the whole run), GPU test suite, smoke, default bench, one short line per workload
mkdir -p gpurun_out
exec > >(tee gpurun_out/status.log) 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
S='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["roofline"]["frac"], d["ms_per_step"], d["roofline"]["kernel"], d.get("verify",{}).get("ok") if d.get("verify") else None)'
echo "=== sanity bench"; timeout -s ABRT 150 python -X faulthandler bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e | python -c "$S" || { echo "SANITY BENCH FAILED"; exit 1; }
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -8
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for wl in d3q19_srt_fp16s_512 d3q19_srt_fp32_256 d3q19_srt_fp32_512 d3q19_srt_fp16s_256 d3q19_srt_fp16c_512 d3q27_trt_fp32_windtunnel_full d3q19_srt_fp32_512_subgrid d3q19_srt_fp16s_512_subgrid d3q19_srt_fp32_256_cavity_mb; do
  echo "=== bench $wl"; timeout -s ABRT 200 python -X faulthandler bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline --no-e2e | tee gpurun_out/status_$wl.json | python -c "$S"
done
for wl in d3q19_srt_fp32_256 d3q19_srt_fp32_512; do
  echo "=== bench $wl variant 32 (occupancy form)"; timeout -s ABRT 200 python -X faulthandler bench.py --workload $wl --variant 32 --steps 100 --warmup 10 --no-cpu-baseline --no-e2e | python -c "$S"
done
echo "=== bench (default, full)"; timeout 600 python bench.py | tee gpurun_out/bench_default.json | cut -c1-2500
