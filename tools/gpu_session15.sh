#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session15.log) 2>&1
for ax in 0 1 2; do timeout 300 python tools/xhalo_probe.py $ax fp16s; done
timeout 300 python tools/xhalo_probe.py 0 fp32
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_pipe -s 6 -c 1 -o gpurun_out/prof15_xhalo_fp16s python tools/xhalo_probe.py 0 fp16s 6 > gpurun_out/ncu15.log 2>&1
