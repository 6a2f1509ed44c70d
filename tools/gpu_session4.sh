#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/session4.log) 2>&1
python tools/debug_parity.py
