#!/usr/bin/env python3
"""Build oracle/_ref/opencl/: the UNMODIFIED reference program (FluidX3D v3.7) as OpenCL executables, one per defines.hpp
configuration, to run beside this repository's CUDA path on a GPU box whose driver ships an OpenCL ICD.

TEST INFRASTRUCTURE ONLY. Runs only where the reference tree is mounted (this container); the GPU box uses the prebuilt
binaries (oracle/_ref/ is git-ignored but travels with gpurun). No reference source is copied into the repository: the tree
is copied to a scratch directory under /tmp, src/defines.hpp is switched per configuration there (comment / uncomment of its
own #define lines, nothing else), and it is compiled with the reference's own command line (make.sh:26).

  FluidX3D_bench_<fp32|fp16s|fp16c>    the reference's own BENCHMARK setup (src/setup.cpp:5-36), untouched: prints "Peak MLUPs/s"
  FluidX3D_<variant>                   src/setup.cpp replaced by tests/scenes/file_scene.cpp (file in, N steps, file out / timing)
  FluidX3D_<variant>+ptx               the same, compiled with the reference's own -DPTX switch (src/opencl.hpp:328-330): the program also
                                       writes the driver-generated PTX of its kernels to ./bin/kernel.ptx (read with ptxas -v / cuobjdump)

Variant names as in build_ref.py: q<19|27>_<srt|trt>_<fp32|fp16s|fp16c>_f<mask> (bit0 VOLUME_FORCE, bit1 EQUILIBRIUM_BOUNDARIES,
bit3 SUBGRID, bit4 MOVING_BOUNDARIES).
"""
import argparse, os, re, shutil, subprocess, sys, tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref", "opencl")
CXX = "/usr/bin/g++"
SCENE = os.path.join(os.path.dirname(os.path.dirname(HERE)), "tests", "scenes", "file_scene.cpp")
DEFAULT = ["bench_fp32", "bench_fp16s", "bench_fp16c",
           "q19_srt_fp32_f0", "q19_srt_fp16s_f0", "q19_srt_fp16c_f0", "q19_trt_fp16s_f3", "q27_trt_fp32_f3", "q27_srt_fp16c_f0",
           "q19_srt_fp32_f16", "q19_srt_fp16s_f8", "q19_srt_fp32_f0+ptx", "q19_srt_fp16s_f0+ptx", "q19_srt_fp16c_f0+ptx"]
SWITCHES = ["D2Q9", "D3Q15", "D3Q19", "D3Q27", "SRT", "TRT", "FP16S", "FP16C", "BENCHMARK", "VOLUME_FORCE", "FORCE_FIELD", "EQUILIBRIUM_BOUNDARIES",
            "MOVING_BOUNDARIES", "SURFACE", "TEMPERATURE", "SUBGRID", "PARTICLES", "INTERACTIVE_GRAPHICS", "INTERACTIVE_GRAPHICS_ASCII", "GRAPHICS"]


def wanted_defines(name):
    name = name.split("+")[0]
    if name.startswith("bench_"):
        st = name.split("_")[1]
        on = {"D3Q19", "SRT", "BENCHMARK"}
    else:
        q, coll, st, f = name.split("_")
        mask = int(f[1:])
        on = {"D3Q" + q[1:], coll.upper()}
        for bit, d in ((1, "VOLUME_FORCE"), (2, "EQUILIBRIUM_BOUNDARIES"), (8, "SUBGRID"), (16, "MOVING_BOUNDARIES")):
            if mask & bit: on.add(d)
    if st != "fp32": on.add(st.upper())
    return on


def switch_defines(text, on):
    """comment / uncomment the top-level switch lines of src/defines.hpp (lines 5-29); everything else stays as it is"""
    out = []
    for line in text.split("\n"):
        m = re.match(r"^(//)?#define (\w+)( //.*)?$", line)
        if m and m.group(2) in SWITCHES:
            line = ("" if m.group(2) in on else "//") + "#define " + m.group(2) + (m.group(3) or "")
        out.append(line)
    return "\n".join(out)


def build(name, ref, scratch):
    src = os.path.join(scratch, name, "src")
    shutil.copytree(os.path.join(ref, "src"), src)
    os.chmod(src, 0o755)
    dpath = os.path.join(src, "defines.hpp")
    os.chmod(dpath, 0o644)
    open(dpath, "w").write(switch_defines(open(os.path.join(ref, "src", "defines.hpp")).read(), wanted_defines(name)))
    if not name.startswith("bench_"):
        spath = os.path.join(src, "setup.cpp")
        os.chmod(spath, 0o644)
        shutil.copyfile(SCENE, spath)
    objs = []
    cpps = sorted(f for f in os.listdir(src) if f.endswith(".cpp"))
    jobs = []
    for f in cpps:
        o = os.path.join(scratch, name, f[:-4] + ".o")
        objs.append(o)
        jobs.append([CXX, "-c", os.path.join(src, f), "-o", o, "-std=c++17", "-pthread", "-O", "-Wno-comment", "-w", "-I" + os.path.join(src, "OpenCL", "include")]
                    + (["-DPTX"] if name.endswith("+ptx") else []))
    return name, src, objs, jobs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--variants", default=",".join(DEFAULT))
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    stamp = lambda n: os.path.join(OUT, "FluidX3D_" + n)
    newest_input = max(os.path.getmtime(SCENE), os.path.getmtime(os.path.join(HERE, "build_opencl_ref.py")))
    names = [n for n in args.variants.split(",") if n and (args.force or not os.path.exists(stamp(n)) or os.path.getmtime(stamp(n)) < newest_input)]
    if not names:
        print("oracle/_ref/opencl is up to date"); return 0
    scratch = tempfile.mkdtemp(prefix="fx3d_opencl_ref_")
    try:
        plans = [build(n, args.reference, scratch) for n in names]
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            results = list(ex.map(lambda j: subprocess.run(j, capture_output=True, text=True), [j for p in plans for j in p[3]]))
        bad = [r for r in results if r.returncode != 0]
        if bad:
            sys.stderr.write(bad[0].stderr[-3000:]); return 1
        for name, src, objs, _ in plans:
            r = subprocess.run([CXX, *objs, "-o", stamp(name), "-pthread", "-L" + os.path.join(src, "OpenCL", "lib"), "-lOpenCL"], capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stderr[-3000:]); return 1
            print("built", stamp(name))
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
