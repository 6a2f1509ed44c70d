#!/bin/bash
# usage: gpu_multi.sh N  -- multi-GPU correctness (bit-identical to the oracle) and weak-scaling bench on N GPUs of one box
N=${1:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/multi_$N.log) 2>&1
nvidia-smi topo -m | head -12
echo "=== dist_check N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py 2>&1 | grep -v "^W\|^\[W\|UserWarning\|warnings.warn" | tail -20
echo "=== bench N=1"
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e | tee gpurun_out/scale_1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['roofline']['frac'], d['ms_per_step'])"
echo "=== bench N=2 split along x (staged faces)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --split 2,1,1 --steps 100 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['config']['domains'], d['value'], d['roofline']['frac'], d['ms_per_step'])"
echo "=== bench N=2 split along y"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --split 1,2,1 --steps 100 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['config']['domains'], d['value'], d['roofline']['frac'], d['ms_per_step'])"
for n in 2 4 8; do
  if [ $n -le $N ]; then
    echo "=== bench N=$n"
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | grep '^{' | tee gpurun_out/scale_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['roofline']['frac'], d['ms_per_step'], d['e2e'])"
  fi
done
