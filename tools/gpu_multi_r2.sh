#!/bin/bash
# round-2 multi-GPU run on N GPUs of one box: bit-identity check on the real peer links, then the bench lines (reference split convention; for N=2 also the z-stacked split)
N=${1:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/multi_r2_$N.log) 2>&1
echo "=== dist_check"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py 2>&1 | grep "DIST_CHECK\|MISMATCH\|rror"
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline "${@:2}" 2>&1 | grep '^{' | tee gpurun_out/scale_r2_${N}_$1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['config']['workload'], d['config']['domains'], d['roofline']['kernel'], d['value'], d['ms_per_step'], 'verify', d['verify']['ok'] if d.get('verify') else None, 'e2e', d['e2e']['value'] if d.get('e2e') else None)"; }
run 29621
run 29622 --workload d3q19_srt_fp32_512 --no-e2e
if [ "$N" = "2" ]; then run 29623 --split 1,1,2 --no-e2e; fi
if [ "$N" = "8" ]; then run 29624 --split 1,2,4 --no-e2e; fi
