#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session27.log) 2>&1
echo "=== parity"; timeout 900 python -m pytest tests -m gpu -q -x -k "segment or hybrid or decomposed" 2>&1 | tail -3
for ax in 0 2; do timeout 300 python tools/xhalo_probe.py $ax fp16s; done
timeout 300 python tools/xhalo_probe.py 0 fp32
timeout 300 python tools/xhalo_probe.py 0 fp16c
echo "=== 1024^3 fp16c"; timeout 600 python bench.py --workload d3q19_srt_fp16c_1024 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"
