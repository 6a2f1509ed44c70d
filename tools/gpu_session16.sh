#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session16.log) 2>&1
for ax in 0 2; do timeout 300 python tools/xhalo_probe.py $ax fp16s; done
timeout 300 python tools/xhalo_probe.py 0 fp32
timeout 300 python tools/xhalo_probe.py 0 fp16c
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "=== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -3 gpurun_out/launches_r01_final.csv
