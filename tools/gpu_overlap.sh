#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/overlap_$N.log) 2>&1
echo "=== dist_check"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py 2>&1 | grep "DIST_CHECK\|MISMATCH\|rror"
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline --no-e2e "${@:2}" 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['config']['domains'], d['config'].get('overlap'), d['value'], d['ms_per_step'])"; }
echo "=== no overlap"; run 29621 --overlap 0
for r in 0 8 16 32; do echo "=== overlap reserve $r"; run 2963$((r%10)) --overlap 1 --reserve $r; done
echo "=== x split: no overlap / overlap"; run 29641 --overlap 0 --split $N,1,1; run 29642 --overlap 1 --split $N,1,1
echo "=== fp32: no overlap / overlap"; run 29643 --overlap 0 --workload d3q19_srt_fp32_512; run 29644 --overlap 1 --workload d3q19_srt_fp32_512
