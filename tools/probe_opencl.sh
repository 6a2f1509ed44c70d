#!/bin/bash
# Probe a GPU box for an OpenCL ICD (NVIDIA's driver normally ships libnvidia-opencl.so.1); writes gpurun_out/opencl_probe.log
out=gpurun_out/opencl_probe.log; mkdir -p gpurun_out; exec > $out 2>&1
echo "== /etc/OpenCL"; ls -laR /etc/OpenCL 2>&1
echo "== env"; env | grep -i -E "nvidia|opencl|ocl" 
echo "== ldconfig"; ldconfig -p | grep -i -E "opencl|nvidia-opencl|nvidia-ptxjit|libcuda|nvidia-nvvm"
echo "== find"; find / -xdev \( -name "libnvidia-opencl*" -o -name "*.icd" -o -name "libOpenCL*" -o -name "clinfo" -o -name "libnvidia-nvvm*" -o -name "libpocl*" \) 2>/dev/null
echo "== nvidia-smi"; nvidia-smi -L; nproc; lscpu | head -20
echo "== ctypes"
python - <<'PY'
import ctypes, glob, os
cands = ["libOpenCL.so.1", "libOpenCL.so"] + glob.glob("/usr/local/cuda*/**/libOpenCL.so*", recursive=True) + glob.glob("/usr/lib/x86_64-linux-gnu/libnvidia-opencl.so*") + glob.glob("/usr/lib64/libnvidia-opencl.so*")
for c in cands:
    try:
        L = ctypes.CDLL(c)
    except OSError as e:
        print(c, "dlopen failed:", e); continue
    n = ctypes.c_uint(0)
    try:
        r = L.clGetPlatformIDs(0, None, ctypes.byref(n))
        print(c, "clGetPlatformIDs ->", r, "platforms:", n.value)
        if r == 0 and n.value:
            ids = (ctypes.c_void_p * n.value)()
            L.clGetPlatformIDs(n.value, ids, None)
            for p in ids:
                buf = ctypes.create_string_buffer(256)
                for what in (0x0902, 0x0903, 0x0901):
                    L.clGetPlatformInfo(ctypes.c_void_p(p), what, 256, buf, None); print("   ", buf.value)
                nd = ctypes.c_uint(0)
                r = L.clGetDeviceIDs(ctypes.c_void_p(p), ctypes.c_ulong(0xFFFFFFFF), 0, None, ctypes.byref(nd)); print("    devices:", r, nd.value)
    except AttributeError as e:
        print(c, "no symbol:", e)
PY
