#!/bin/bash
# one ncu --set full capture of the dominant kernel of a workload (one launch). usage: gpu_ncu.sh <workload> <out name> [kernel regex] [extra bench args]
mkdir -p gpurun_out
wl=$1; out=$2; k=${3:-k_stream_collide}; shift 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/$out python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-verify "$@" > gpurun_out/$out.log 2>&1
tail -3 gpurun_out/$out.log; ls -la gpurun_out/$out.ncu-rep
