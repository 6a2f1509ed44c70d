#!/bin/bash
# Tuning builds: recompile ONE (velocity set, storage) translation unit of sc_inst.cu with extra defines and link it with the objects of the
# regular build into fluidx3d_b200/libfx3d_cuda_<tag>.so (picked up with FX3D_LIB=... python bench.py). usage: build_variant.sh <tag> <Q> <ST> <defines...>
set -e
tag=$1; Q=$2; ST=$3; shift 3
cd "$(dirname "$0")/../fluidx3d_b200/csrc"
OBJ=../../build/csrc
NVFLAGS="-gencode arch=compute_100a,code=sm_100a -diag-suppress 550 -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr -Xptxas -v -DFX3D_V4_MINBLOCKS=3 -DFX3D_V2_MINBLOCKS=4 -DFX3D_PIPE_MINBLOCKS=4 -DFX3D_FUSED_COLLIDE=0 -DFX3D_PIPE_STAGES=2"
/usr/local/cuda/bin/nvcc $NVFLAGS "$@" -DFX3D_Q=$Q -DFX3D_ST=$ST -c sc_inst.cu -o $OBJ/sc_${Q}_${ST}_$tag.o 2> $OBJ/sc_${Q}_${ST}_$tag.ptxas.log
objs=""
for c in 19_0 19_1 19_2 27_0 27_1 27_2; do if [ "$c" = "${Q}_${ST}" ]; then objs="$objs $OBJ/sc_${c}_$tag.o"; else objs="$objs $OBJ/sc_$c.o"; fi; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libfx3d_cuda_$tag.so $objs $OBJ/fx3d_lbm.o $OBJ/fx3d_runtime.o -lcudart
echo "built fluidx3d_b200/libfx3d_cuda_$tag.so"
