#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session17.log) 2>&1
for r in 64 128 256 512 1024; do echo "px_round $r"; FX3D_PX_ROUND=$r timeout 300 python tools/xhalo_probe.py 0 fp16s; done
for r in 128 256; do echo "px_round $r (no x halo)"; FX3D_PX_ROUND=$r timeout 300 python tools/xhalo_probe.py 2 fp16s; done
