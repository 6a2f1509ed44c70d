// ocl_shim.hpp -- the handful of OpenCL C built-ins and types the reference's LBM device functions use,
// provided as plain C++ so that the reference's own kernel source (extracted at build time from
// /root/reference/src/kernel.cpp, never committed) can be compiled and run natively as oracle/_ref.
// TEST INFRASTRUCTURE ONLY (see oracle/lbm_oracle.h).
//
// Semantics: every built-in here is the exactly-rounded IEEE operation the OpenCL C 1.2 specification
// defines for it (fma: single rounding; clamp: fmin(fmax(x,lo),hi); vstore_half_rte: binary16 round to
// nearest even; vload_half: exact widening). Compile with -ffp-contract=off.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>

namespace ocl {

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long ulong; // 64 bit on LP64
typedef _Float16 half;

struct uint3 { uint x, y, z; };
struct float3 { float x, y, z; };
static inline uint3 make_uint3(uint x, uint y, uint z) { return uint3{ x, y, z }; }
static inline float3 make_float3(float x, float y, float z) { return float3{ x, y, z }; }

// float3 arithmetic of the voxeliser (src/kernel.cpp:2267-2357): component-wise, every operation separately rounded; cross() and dot() as the
// plain expressions (the same ones the reference's host-side float3 uses, src/utilities.hpp:1020-1025)
static inline float3 operator+(const float3 a, const float3 b) { return float3{ a.x+b.x, a.y+b.y, a.z+b.z }; }
static inline float3 operator-(const float3 a, const float3 b) { return float3{ a.x-b.x, a.y-b.y, a.z-b.z }; }
static inline float3 operator*(const float s, const float3 a) { return float3{ s*a.x, s*a.y, s*a.z }; } // scalar times vector (update_force_field, src/kernel.cpp:1883)
static inline float3 cross(const float3 a, const float3 b) { return float3{ a.y*b.z-a.z*b.y, a.z*b.x-a.x*b.z, a.x*b.y-a.y*b.x }; }
static inline float dot(const float3 a, const float3 b) { return a.x*b.x+a.y*b.y+a.z*b.z; }
static inline int clamp(const int x, const int lo, const int hi) { return x<lo ? lo : x>hi ? hi : x; }
static inline uint min(const uint a, const uint b) { return a<b ? a : b; }

static inline uint as_uint(const float x) { uint u; std::memcpy(&u, &x, 4); return u; }
static inline float as_float(const uint u) { float x; std::memcpy(&x, &u, 4); return x; }
static inline float fma(const float a, const float b, const float c) { return __builtin_fmaf(a, b, c); }
static inline float sqrt(const float x) { return __builtin_sqrtf(x); } // correctly rounded, as OpenCL's sqrt on an IEEE device without fast-math (SUBGRID only)
static inline float clamp(const float x, const float lo, const float hi) { return __builtin_fminf(__builtin_fmaxf(x, lo), hi); }
static inline float vload_half(const ulong offset, const half* p) { return (float)p[offset]; }
static inline void vstore_half_rte(const float x, const ulong offset, half* p) { p[offset] = (half)x; }

// run-time stand-ins for the constants LBM_Domain::device_defines() bakes into the JIT source (src/lbm.cpp:334-425)
static uint g_Nx=1u, g_Ny=1u, g_Nz=1u, g_Dx=1u, g_Dy=1u, g_Dz=1u;
static int g_Ox=0, g_Oy=0, g_Oz=0; // def_Ox.. : offset of the domain in the global grid
static ulong g_N=1ul;
static float g_w=1.0f;
static thread_local ulong g_gid=0ul;
static inline ulong get_global_id(const uint) { return g_gid; }

} // namespace ocl
