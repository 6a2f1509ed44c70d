#!/bin/bash
# parity of the two widenings (SUBGRID, MOVING_BOUNDARIES) on the GPU + the reference's lid-driven cavity in its own formulation
mkdir -p gpurun_out
exec > >(tee gpurun_out/next_rows.log) 2>&1
echo "=== parity"; timeout 900 python -m pytest tests -m gpu -q -x -k "subgrid or moving" 2>&1 | tail -3
echo "=== golden"; timeout 600 python -m pytest tests -m gpu -q -x -k "golden" 2>&1 | tail -2
