#!/bin/bash
# Small tuning library: all six (velocity set, storage) translation units compiled with -DFX3D_TUNE_ONLY (only the whole-row SRT kernel of the
# benchmark lines is instantiated: seconds instead of minutes, a few MB instead of 120) plus extra defines, linked with the regular runtime objects
# into fluidx3d_b200/libfx3d_cuda_<tag>.so (FX3D_LIB=... python bench.py). usage: build_tune.sh <tag> <defines...>
set -e
tag=$1; shift
cd "$(dirname "$0")/../fluidx3d_b200/csrc"
OBJ=../../build/tune_$tag; mkdir -p $OBJ
NVFLAGS="-gencode arch=compute_100a,code=sm_100a -diag-suppress 550 -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr -Xptxas -v -DFX3D_V4_MINBLOCKS=3 -DFX3D_V2_MINBLOCKS=4 -DFX3D_PIPE_MINBLOCKS=4 -DFX3D_FUSED_COLLIDE=0 -DFX3D_PIPE_STAGES=2 -DFX3D_TUNE_ONLY"
for c in 19_0 19_1 19_2 27_0 27_1 27_2; do
  /usr/local/cuda/bin/nvcc $NVFLAGS "$@" -DFX3D_Q=${c%_*} -DFX3D_ST=${c#*_} -c sc_inst.cu -o $OBJ/sc_$c.o 2> $OBJ/sc_$c.ptxas.log &
done
LBMO=../../build/csrc/fx3d_lbm.o
if [ -n "$FX3D_TUNE_LBM" ]; then /usr/local/cuda/bin/nvcc $NVFLAGS "$@" -c fx3d_lbm.cu -o $OBJ/fx3d_lbm.o 2> $OBJ/fx3d_lbm.ptxas.log & LBMO=$OBJ/fx3d_lbm.o; fi # (FX3D_TUNE_LBM=1: the defines also reach the host-side layout code)
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libfx3d_cuda_$tag.so $OBJ/sc_*.o $LBMO ../../build/csrc/fx3d_runtime.o -lcudart
for c in 19_0 19_1 19_2 27_1; do echo -n "$c: "; grep -A2 "k_stream_collide_tmaILi${c%_*}ELi0ELi${c#*_}ELb0ELi0E" $OBJ/sc_$c.ptxas.log | grep -E "spill|Used" | tr '\n' ' '; echo; done
ls -la ../libfx3d_cuda_$tag.so
