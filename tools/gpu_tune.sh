#!/bin/bash
# tuning run: one short device-resident bench line per (library build, workload). usage: gpu_tune.sh "<lib tags ('-' = regular build)>" "<workloads>"
mkdir -p gpurun_out
exec > >(tee -a gpurun_out/tune.log) 2>&1
S='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["roofline"]["frac"], d["ms_per_step"], d["roofline"]["kernel"], d.get("verify",{}).get("ok") if d.get("verify") else None)'
for tag in $1; do
  lib=fluidx3d_b200/libfx3d_cuda.so; [ "$tag" != "-" ] && lib=fluidx3d_b200/libfx3d_cuda_$tag.so
  for wl in $2; do
    echo -n "$tag $wl: "; FX3D_LIB=$PWD/$lib timeout -s ABRT 120 python -X faulthandler bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline --no-e2e $3 | python -c "$S"
  done
done
