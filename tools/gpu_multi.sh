#!/bin/bash
# usage: gpu_multi.sh N [list of bench sizes] -- multi-GPU correctness (bit-identical to the oracle) and weak-scaling bench on N GPUs of one box
N=${1:-2}
SIZES=${2:-"1 2 4 8"}
mkdir -p gpurun_out
exec > >(tee gpurun_out/multi_$N.log) 2>&1
nvidia-smi topo -m | head -12
echo "=== dist_check N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py 2>&1 | grep "dist_check\|DIST_CHECK"
for n in $SIZES; do
  if [ $n -le $N ]; then
    echo "=== bench N=$n"
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline | tee gpurun_out/scale_1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['roofline']['frac'], d['ms_per_step'], d['e2e']['value'])"
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | grep '^{' | tee gpurun_out/scale_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['config']['domains'], d['value'], d['roofline']['frac'], d['ms_per_step'], d['e2e']['value'])"
    fi
  fi
done
