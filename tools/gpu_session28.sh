#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session28.log) 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_tma_seg -s 6 -c 1 -o gpurun_out/prof28_xhalo_seg python tools/xhalo_probe.py 0 fp16s 6 > gpurun_out/ncu28.log 2>&1
tail -2 gpurun_out/ncu28.log | cut -c1-200
