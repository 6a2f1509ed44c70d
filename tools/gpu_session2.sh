#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session2.log) 2>&1
echo "=== microbench"; ./tools/microbench/ubench
echo "=== pytest gpu (all)"; time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30
