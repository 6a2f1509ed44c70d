#!/usr/bin/env python3
"""Build oracle/_ref: the reference's OWN LBM device code, compiled natively for the host CPU.

TEST INFRASTRUCTURE ONLY. Runs only where the reference tree is mounted (this container); the GPU box
uses the prebuilt oracle/_ref/*.so. Nothing from the reference is copied into the repository: the
OpenCL C program text is produced at build time by the reference's own get_opencl_c_code()
(src/kernel.hpp:7-18, src/kernel.cpp) into oracle/_ref/ (git-ignored), preprocessed per variant with a
define prologue equivalent to LBM_Domain::device_defines() (src/lbm.cpp:334-468, grid constants redirected
to run-time variables), the hot-path functions are cut out by name and wrapped with oracle/ref/ocl_shim.hpp
plus a tiny NDRange driver.

usage: build_ref.py [--reference /root/reference] [--variants q19_srt_fp32_f0,...]
Variant name: q<19|27>_<srt|trt>_<fp32|fp16s|fp16c>_f<mask>  (mask bit0 VOLUME_FORCE, bit1 EQUILIBRIUM_BOUNDARIES,
bit2 UPDATE_FIELDS, bit3 SUBGRID, bit4 MOVING_BOUNDARIES, bit5 FORCE_FIELD)
"""
import argparse, os, re, subprocess, sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
CXX = "/usr/bin/g++"

DEFAULT_VARIANTS = [
    "q19_srt_fp32_f0", "q19_srt_fp16s_f0", "q19_srt_fp16c_f0",
    "q19_trt_fp32_f0", "q19_srt_fp32_f1", "q19_srt_fp32_f2", "q19_trt_fp16s_f3", "q19_srt_fp32_f4",
    "q27_srt_fp32_f0", "q27_trt_fp32_f3", "q27_srt_fp16s_f0", "q27_trt_fp16c_f3",
    "q19_srt_fp32_f8", "q19_trt_fp16s_f11", "q27_srt_fp16c_f8", "q19_srt_fp16s_f8",  # bit3: SUBGRID (first "next" row of SURVEY 8f)
    "q19_srt_fp32_f16", "q19_trt_fp16s_f19", "q27_srt_fp16c_f18", "q19_srt_fp16s_f24",  # bit4: MOVING_BOUNDARIES
    "q19_srt_fp32_f33", "q19_trt_fp16s_f35", "q27_srt_fp16c_f33", "q19_srt_fp32_f34",  # bit5: FORCE_FIELD (with and without VOLUME_FORCE)
]

# functions on the hot path; everything else in the program text is dropped
WANTED = [
    "sq", "coordinates", "index", "is_halo", "half_to_float_custom", "float_to_half_custom", "index_f", "c", "w",
    "calculate_indices", "neighbors", "load3", "store3", "calculate_f_eq", "calculate_rho_u", "calculate_forcing_terms",
    "apply_moving_boundaries", "load_f", "store_f", "initialize", "update_moving_boundaries", "stream_collide", "update_fields",
    "get_area", "index_extract_p", "index_extract_m", "index_insert_p", "index_insert_m", "index_transfer",
    "extract_fi", "insert_fi", "transfer_extract_fi", "transfer__insert_fi",
    "extract_rho_u_flags", "insert_rho_u_flags", "transfer_extract_rho_u_flags", "transfer__insert_rho_u_flags",
    "position", "voxelize_mesh", "unvoxelize_mesh",
    "update_force_field", "reset_force_field", "extract_F", "insert_F", "transfer_extract_F", "transfer__insert_F",  # SURVEY 8f rank 4 (FORCE_FIELD builds only)
]
FORCE_FIELD_ONLY = ("update_force_field", "reset_force_field", "extract_F", "insert_F", "transfer_extract_F", "transfer__insert_F")


def prologue(q, coll, storage, mask):
    transfers = 5 if q == 19 else 9
    d = [
        "#define def_Nx g_Nx", "#define def_Ny g_Ny", "#define def_Nz g_Nz", "#define def_N g_N", "#define uxx uint",
        "#define def_Dx g_Dx", "#define def_Dy g_Dy", "#define def_Dz g_Dz", "#define def_Ox g_Ox", "#define def_Oy g_Oy", "#define def_Oz g_Oz",
        "#define def_Ax (g_Ny*g_Nz)", "#define def_Ay (g_Nz*g_Nx)", "#define def_Az (g_Nx*g_Ny)",
        f"#define D3Q{q}", f"#define def_velocity_set {q}u", "#define def_dimensions 3u", f"#define def_transfers {transfers}u",
        "#define def_c 0.57735027f", "#define def_w g_w",
    ]
    if q == 19:
        d += ["#define def_w0 (1.0f/3.0f)", "#define def_ws (1.0f/18.0f)", "#define def_we (1.0f/36.0f)"]
    else:
        d += ["#define def_w0 (1.0f/3.375f)", "#define def_ws (1.0f/13.5f)", "#define def_we (1.0f/54.0f)", "#define def_wc (1.0f/216.0f)"]
    d += [f"#define {coll.upper()}"]
    d += ["#define TYPE_S 0x01", "#define TYPE_E 0x02", "#define TYPE_T 0x04", "#define TYPE_F 0x08", "#define TYPE_I 0x10",
          "#define TYPE_G 0x20", "#define TYPE_X 0x40", "#define TYPE_Y 0x80", "#define TYPE_MS 0x03", "#define TYPE_BO 0x03",
          "#define TYPE_IF 0x18", "#define TYPE_IG 0x30", "#define TYPE_GI 0x38", "#define TYPE_SU 0x38", "#define TYPE_XY 0xC0"]
    if storage == "fp16s":
        d += ["#define fpxx half", "#define fpxx_copy ushort", "#define load(p,o) (vload_half(o,p)*3.0517578E-5f)",
              "#define store(p,o,x) vstore_half_rte((x)*32768.0f,o,p)"]
    elif storage == "fp16c":
        d += ["#define fpxx ushort", "#define fpxx_copy ushort", "#define load(p,o) half_to_float_custom((p)[o])",
              "#define store(p,o,x) (p)[o]=float_to_half_custom(x)"]
    else:
        d += ["#define fpxx float", "#define fpxx_copy float", "#define load(p,o) (p)[o]", "#define store(p,o,x) (p)[o]=(x)"]
    if mask & 1: d.append("#define VOLUME_FORCE")
    if mask & 2: d.append("#define EQUILIBRIUM_BOUNDARIES")
    if mask & 4: d.append("#define UPDATE_FIELDS")
    if mask & 8: d.append("#define SUBGRID")
    if mask & 16: d.append("#define MOVING_BOUNDARIES")
    if mask & 32: d.append("#define FORCE_FIELD")
    return "\n".join(d) + "\n"


def top_level_functions(text):
    """split preprocessed program text into top-level brace-balanced chunks -> {name: source}"""
    out, depth, start = {}, 0, 0
    for pos, ch in enumerate(text):
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                chunk = text[start:pos + 1].strip()
                start = pos + 1
                m = re.search(r"(\w+)\s*\(", chunk)
                if m:
                    out[m.group(1)] = chunk
    return out


DRIVER = r"""
extern "C" {
using namespace ocl;
static int g_threads = 0;
void ref_set_threads(int n) { g_threads = n; }
int ref_get_threads() { return g_threads>0 ? g_threads : omp_get_max_threads(); }
void ref_set_grid(uint Nx, uint Ny, uint Nz, uint Dx, uint Dy, uint Dz) {
	g_Nx = Nx; g_Ny = Ny; g_Nz = Nz; g_Dx = Dx; g_Dy = Dy; g_Dz = Dz; g_N = (ulong)Nx*(ulong)Ny*(ulong)Nz;
}
void ref_set_w(float w) { g_w = w; }
uint ref_velocity_set() { return def_velocity_set; }
uint ref_bytes_per_ddf() { return (uint)sizeof(fpxx); }
#define NDRANGE(range, call) _Pragma("omp parallel for schedule(static) num_threads(ref_get_threads())") \
	for(ulong gid=0ul; gid<(ulong)(range); gid++) { g_gid = gid; call; }
void ref_initialize(void* fi, const float* rho, float* u, uchar* flags) { NDRANGE(g_N, initialize((fpxx*)fi, rho, u, flags)) }
#ifndef FORCE_FIELD
void ref_stream_collide(void* fi, float* rho, float* u, uchar* flags, ulong t, float fx, float fy, float fz) { NDRANGE(g_N, stream_collide((fpxx*)fi, rho, u, flags, t, fx, fy, fz)) }
void ref_update_fields(const void* fi, float* rho, float* u, const uchar* flags, ulong t, float fx, float fy, float fz) { NDRANGE(g_N, update_fields((const fpxx*)fi, rho, u, flags, t, fx, fy, fz)) }
#else // FORCE_FIELD: the kernels take the force field as an extra argument (src/kernel.cpp:1455-1457,1795-1797); the object_* reductions are left out
// (work-group local memory and barriers; their floating-point atomics make the reference's own result order-dependent anyway)
void ref_stream_collide_F(void* fi, float* rho, float* u, uchar* flags, ulong t, float fx, float fy, float fz, const float* F) { NDRANGE(g_N, stream_collide((fpxx*)fi, rho, u, flags, t, fx, fy, fz, F)) }
void ref_update_fields_F(const void* fi, float* rho, float* u, const uchar* flags, ulong t, float fx, float fy, float fz, const float* F) { NDRANGE(g_N, update_fields((const fpxx*)fi, rho, u, flags, t, fx, fy, fz, F)) }
void ref_update_force_field(const void* fi, const uchar* flags, ulong t, float* F) { NDRANGE(g_N, update_force_field((const fpxx*)fi, flags, t, F)) }
void ref_reset_force_field(float* F) { NDRANGE(g_N, reset_force_field(F)) }
void ref_transfer_extract_F(uint direction, ulong t, void* bp, void* bm, const float* F) { NDRANGE(get_area(direction), transfer_extract_F(direction, t, (float*)bp, (float*)bm, F)) }
void ref_transfer_insert_F(uint direction, ulong t, const void* bp, const void* bm, float* F) { NDRANGE(get_area(direction), transfer__insert_F(direction, t, (const float*)bp, (const float*)bm, F)) }
#endif
void ref_transfer_extract_fi(uint direction, ulong t, void* bp, void* bm, const void* fi) { NDRANGE(get_area(direction), transfer_extract_fi(direction, t, (fpxx_copy*)bp, (fpxx_copy*)bm, (const fpxx_copy*)fi)) }
void ref_transfer_insert_fi(uint direction, ulong t, const void* bp, const void* bm, void* fi) { NDRANGE(get_area(direction), transfer__insert_fi(direction, t, (const fpxx_copy*)bp, (const fpxx_copy*)bm, (fpxx_copy*)fi)) }
void ref_transfer_extract_rho_u_flags(uint direction, ulong t, void* bp, void* bm, const float* rho, const float* u, const uchar* flags) { NDRANGE(get_area(direction), transfer_extract_rho_u_flags(direction, t, (char*)bp, (char*)bm, rho, u, flags)) }
void ref_transfer_insert_rho_u_flags(uint direction, ulong t, const void* bp, const void* bm, float* rho, float* u, uchar* flags) { NDRANGE(get_area(direction), transfer__insert_rho_u_flags(direction, t, (const char*)bp, (const char*)bm, rho, u, flags)) }
#ifdef MOVING_BOUNDARIES
void ref_update_moving_boundaries(const float* u, uchar* flags) { NDRANGE(g_N, update_moving_boundaries(u, flags)) }
#endif
void ref_set_offsets(int Ox, int Oy, int Oz) { g_Ox = Ox; g_Oy = Oy; g_Oz = Oz; }
void ref_voxelize_mesh(uint direction, void* fi, float* u, uchar* flags, ulong t, uchar flag, const float* p0, const float* p1, const float* p2, const float* bbu) {
	NDRANGE(get_area(direction), voxelize_mesh(direction, (fpxx*)fi, u, flags, t, flag, p0, p1, p2, bbu)) }
void ref_unvoxelize_mesh(uchar* flags, uchar flag, float x0, float y0, float z0, float x1, float y1, float z1) { NDRANGE(g_N, unvoxelize_mesh(flags, flag, x0, y0, z0, x1, y1, z1)) }
ushort ref_float_to_half_custom(float x) { return float_to_half_custom(x); }
float ref_half_to_float_custom(ushort x) { return half_to_float_custom(x); }
}
"""


def build_variant(name, cl_path):
    m = re.fullmatch(r"q(19|27)_(srt|trt)_(fp32|fp16s|fp16c)_f(\d+)", name)
    if not m:
        raise SystemExit(f"bad variant name {name}")
    q, coll, storage, mask = int(m.group(1)), m.group(2), m.group(3), int(m.group(4))
    src = prologue(q, coll, storage, mask) + open(cl_path).read()
    pre = subprocess.run(["/usr/bin/cpp", "-P", "-undef", "-nostdinc", "-x", "c", "-"], input=src, capture_output=True, text=True, check=True).stdout
    pre = "\n".join(l for l in pre.splitlines() if not l.lstrip().startswith("#"))
    text = " ".join(pre.split())
    funcs = top_level_functions(text)
    body = []
    for fn in WANTED:
        if fn == "calculate_forcing_terms" and not (mask & 1):
            continue
        if fn in ("apply_moving_boundaries", "update_moving_boundaries") and not (mask & 16):
            continue
        if fn in FORCE_FIELD_ONLY and not (mask & 32):
            continue
        if fn not in funcs:
            raise SystemExit(f"{name}: function {fn} not found in the reference program text")
        s = funcs[fn]
        s = s.replace("(uint3)(", "make_uint3(").replace("(float3)(", "make_float3(")
        s = re.sub(r"\bkernel\b", "", s)
        s = re.sub(r"\bglobal\b", "", s)
        body.append(s)
    gen = os.path.join(OUT, f"ref_{name}.cpp")
    defs = prologue(q, coll, storage, mask)  # the driver needs def_velocity_set / fpxx too
    with open(gen, "w") as f:
        f.write("// GENERATED at build time from the mounted reference tree -- do not commit.\n")
        f.write('#include "ocl_shim.hpp"\n#include <omp.h>\n' + defs + "namespace ocl {\n" + "\n".join(body) + "\n}\n" + DRIVER)
    so = os.path.join(OUT, f"libref_{name}.so")
    cmd = [CXX, "-std=c++17", "-O3", "-mavx2", "-mfma", "-mf16c", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared", "-w",
           "-I", HERE, "-o", so, gen]
    subprocess.run(cmd, check=True)
    return so


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--variants", default=",".join(DEFAULT_VARIANTS))
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(a.reference, "src")):
        print(f"reference tree not found at {a.reference}; keeping any prebuilt oracle/_ref", file=sys.stderr)
        return 0
    os.makedirs(OUT, exist_ok=True)
    gen_cl, cl = os.path.join(OUT, "gen_cl"), os.path.join(OUT, "kernel.cl")
    subprocess.run([CXX, "-std=c++17", "-O0", "-w", "-I", os.path.join(a.reference, "src"), "-o", gen_cl,
                    os.path.join(HERE, "gen_main.cpp"), os.path.join(a.reference, "src", "kernel.cpp")], check=True)
    subprocess.run([gen_cl, cl], check=True)
    # a second tiny tool: the reference's own to_string(float) (src/utilities.hpp:2745-2754) for pinning orc_float_to_string
    subprocess.run([CXX, "-std=c++17", "-O1", "-w", "-I", os.path.join(a.reference, "src"), "-o", os.path.join(OUT, "ref_to_string"),
                    os.path.join(HERE, "to_string_main.cpp")], check=True)
    names = [v for v in a.variants.split(",") if v]
    with ThreadPoolExecutor(max_workers=8) as ex:
        for so in ex.map(lambda n: build_variant(n, cl), names):
            print("built", os.path.relpath(so))
    return 0


if __name__ == "__main__":
    sys.exit(main())
