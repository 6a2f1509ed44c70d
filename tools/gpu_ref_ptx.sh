#!/bin/bash
# dump the PTX the NVIDIA OpenCL driver generates for the reference's kernels (its own -DPTX switch), for ptxas -v / cuobjdump here
export OCL_ICD_FILENAMES=/usr/lib/libnvidia-opencl.so.1
cd gpurun_out
for st in fp32 fp16s fp16c; do
  mkdir -p bin; rm -f bin/kernel.ptx
  FX3D_REF_N=256,256,256 FX3D_REF_STEPS=2 ../oracle/_ref/opencl/FluidX3D_q19_srt_${st}_f0+ptx > ref_ptx_$st.log 2>&1
  cp bin/kernel.ptx ref_kernel_$st.ptx 2>/dev/null; ls -la ref_kernel_$st.ptx
done
# can Nsight Compute see OpenCL kernels?
# (Nsight Compute does not see OpenCL kernels: "No kernels were profiled")
