#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/multi_final_$N.log) 2>&1
echo "=== dist_check"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py 2>&1 | grep "DIST_CHECK\|MISMATCH\|rror"
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline --no-e2e "${@:2}" 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['config']['workload'], d['config']['domains'], d['roofline']['kernel'], d['value'], d['ms_per_step'])"; }
run 29621
run 29622 --workload d3q19_srt_fp32_512
run 29623 --workload d3q19_srt_fp32_512 --split $N,1,1
run 29624 --split $N,1,1
run 29625 --workload d3q19_srt_fp32_512 --strong
