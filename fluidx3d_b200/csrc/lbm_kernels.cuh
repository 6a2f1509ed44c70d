// lbm_kernels.cuh -- device kernels of the LBM hot path for sm_100a.
//
// Replaces the reference's OpenCL kernels (FluidX3D v3.7 src/kernel.cpp): stream_collide :1454-1636 with
// load_f/store_f :1326-1339, initialize :1358-1430, update_fields :1794-1870, transfer_extract_fi /
// transfer__insert_fi :2049-2131, transfer_*_rho_u_flags :2133-2158 -- and adds the direct peer (NVLink) halo
// exchange that replaces LBM::communicate_field (src/lbm.cpp:1355-1383).
//
// DDF memory layout (private to this library; DDFs never exist on the host, src/lbm.hpp:38): structure of arrays,
// one array per slot i, element (x,y,z) of slot i at  i*slot + (x+xo) + px*(y + Ny*z)  where px is the row pitch
// rounded up to 64 elements and xo shifts x so that the first non-halo cell of every row starts on a 16-byte
// boundary. Esoteric-Pull slot semantics are exactly the reference's (which slot holds which population at which
// step parity), so solid-cell bounce-back, TYPE_E behaviour and the halo exchange addresses are unchanged.
// rho, u (SoA, 3 planes) and flags keep the reference's plain layout n = x+(y+z*Ny)*Nx: they are what the host API
// reads and writes.
#pragma once
#include "lbm_core.cuh"

#ifndef FX3D_FUSED_COLLIDE
#define FX3D_FUSED_COLLIDE 0 // 1: relax each direction pair as soon as its equilibrium exists (smaller live set)
#endif
#if FX3D_FUSED_COLLIDE
#define FX3D_COLLIDE collide_cell_fused
#else
#define FX3D_COLLIDE collide_cell
#endif
#ifndef FX3D_V4_MINBLOCKS
#define FX3D_V4_MINBLOCKS 3 // resident 128-thread blocks per SM the 4-cell vector kernel is compiled for (register cap 65536/(128*n))
#endif
#ifndef FX3D_V2_MINBLOCKS
#define FX3D_V2_MINBLOCKS 4 // the same for the 2-cell vector kernel
#endif

namespace fx3d {

struct Lattice { // one LBM_Domain as the device sees it (passed by value)
	uint32_t Nx, Ny, Nz; // local size, halo layers included
	uint32_t Hx, Hy, Hz; // 1 where the axis is decomposed (halo layer present), else 0
	uint32_t px, xo;     // DDF row pitch and x offset, in elements
	uint64_t slot;       // elements between consecutive DDF slots
	uint32_t slot32;     // the same, known to fit 32 bits (checked on the host)
	void* fi;
	float* rho;
	float* u;
	uint8_t* flags;
	float w, fx, fy, fz;
	uint32_t odd;        // t&1
	uint32_t eb;         // EQUILIBRIUM_BOUNDARIES enabled
	uint32_t upd;        // UPDATE_FIELDS enabled
	uint32_t mb;         // MOVING_BOUNDARIES enabled (general kernels only)
	float* F;            // FORCE_FIELD: per-cell force [3N] SoA, or null (general kernels, update_fields, the force-field kernels)
};
struct Region { uint32_t g0, g1, y0, y1, z0, z1; }; // x-group range [g0,g1), cell ranges in y and z

FX3D_HD uint64_t cells(const Lattice& L) { return (uint64_t)L.Nx*L.Ny*L.Nz; }
FX3D_HD uint64_t lin(const Lattice& L, uint32_t x, uint32_t y, uint32_t z) { return (uint64_t)x+((uint64_t)y+(uint64_t)z*L.Ny)*L.Nx; }
FX3D_HD uint64_t row(const Lattice& L, uint32_t y, uint32_t z) { return (uint64_t)L.px*((uint64_t)y+(uint64_t)z*L.Ny); }
FX3D_HD uint64_t phys(const Lattice& L, uint32_t x, uint32_t y, uint32_t z) { return row(L, y, z)+(uint64_t)(x+L.xo); }
FX3D_HD uint32_t inc(uint32_t v, uint32_t n) { return v+1u==n ? 0u : v+1u; }
FX3D_HD uint32_t dec(uint32_t v, uint32_t n) { return v==0u ? n-1u : v-1u; }
FX3D_HD char* mad_wide_again(uint32_t a, uint32_t b, char* c) { // the same, but never merged with an earlier identical computation: recomputing an address costs one instruction, keeping it live across the collision costs two registers
#if defined(FX3D_HOST_EMULATION)
	return c+(uint64_t)a*(uint64_t)b;
#else
	unsigned long long r; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"((unsigned long long)c)); return reinterpret_cast<char*>(r);
#endif
}
// typed global-memory accesses for addresses that come out of the asm above (the compiler would otherwise use generic LD/ST)
template<class E> FX3D_HD E load_global(const E* p) {
#if defined(FX3D_HOST_EMULATION)
	return *p;
#else
	if constexpr(sizeof(E)==4) { uint32_t v; asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return __uint_as_float(v); }
	else { unsigned short v; asm volatile("ld.global.b16 %0, [%1];" : "=h"(v) : "l"(p) : "memory"); return (E)v; }
#endif
}
template<class E> FX3D_HD void store_global(E* p, E v) {
#if defined(FX3D_HOST_EMULATION)
	*p = v;
#else
	if constexpr(sizeof(E)==4) asm volatile("st.global.b32 [%0], %1;" :: "l"(p), "r"(__float_as_uint(v)) : "memory");
	else asm volatile("st.global.b16 [%0], %1;" :: "l"(p), "h"((unsigned short)v) : "memory");
#endif
}
template<int E> FX3D_HD uint32_t step(uint32_t v, uint32_t n) { if constexpr(E>0) return inc(v, n); else if constexpr(E<0) return dec(v, n); else return v; }
FX3D_HD uint32_t step_rt(int e, uint32_t v, uint32_t n) { return e>0 ? inc(v, n) : e<0 ? dec(v, n) : v; }

FX3D_HD char* mad_wide(uint32_t a, uint32_t b, char* c) { // c + a*b with the product taken in 64 bits: one IMAD.WIDE.U32
#if defined(FX3D_HOST_EMULATION)
	return c+(uint64_t)a*(uint64_t)b;
#else
	unsigned long long r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"((unsigned long long)c)); return reinterpret_cast<char*>(r);
#endif
}

// ---- 4 consecutive x-elements of one slot, as loaded; cells (0,1) and (2,3) form the two F2 lane pairs ----
struct alignas(16) Raw16 { unsigned long long x, y; };
struct alignas(8) Raw8 { uint32_t x, y; };
template<int ST, int K> struct Pack;
template<> struct Pack<ST_FP32, 4> {
	F2 p[2];
	FX3D_HD void load(const float* q) { const Raw16 t = *reinterpret_cast<const Raw16*>(q); memcpy_bits(p[0], t.x); memcpy_bits(p[1], t.y); }
	FX3D_HD void store(float* q) const { Raw16 t; t.x = bits64(p[0]); t.y = bits64(p[1]); *reinterpret_cast<Raw16*>(q) = t; }
	FX3D_HD void store_tail(float* q) const { q[1] = f2_hi(p[0]); *reinterpret_cast<unsigned long long*>(q+2) = bits64(p[1]); } // all but the first element
	FX3D_HD void store_head(float* q) const { *reinterpret_cast<unsigned long long*>(q) = bits64(p[0]); q[2] = f2_lo(p[1]); } // all but the last element
	FX3D_HD uint32_t first_bits() const { return __float_as_uint(f2_lo(p[0])); }
	FX3D_HD uint32_t last_bits() const { return __float_as_uint(f2_hi(p[1])); }
	// my 4 elements belong at positions x+1..x+4 (up) / x-1..x+2 (down) of a periodic row of W elements (x a multiple of 4)
	FX3D_HD void store_seg_up(float* row, uint32_t x) const { row[x+1u] = f2_lo(p[0]); *reinterpret_cast<unsigned long long*>(row+x+2u) = bits64(make_f2(f2_hi(p[0]), f2_lo(p[1]))); row[x+4u] = f2_hi(p[1]); } // positions x+1..x+4 of a padded segment buffer
	FX3D_HD void store_seg_down(float* row, uint32_t x) const { *(row+x-1) = f2_lo(p[0]); *reinterpret_cast<unsigned long long*>(row+x) = bits64(make_f2(f2_hi(p[0]), f2_lo(p[1]))); row[x+2u] = f2_hi(p[1]); } // positions x-1..x+2
	FX3D_HD void store_row_up(float* row, uint32_t x, uint32_t W) const { row[x+1u] = f2_lo(p[0]); *reinterpret_cast<unsigned long long*>(row+x+2u) = bits64(make_f2(f2_hi(p[0]), f2_lo(p[1]))); row[x+4u==W ? 0u : x+4u] = f2_hi(p[1]); }
	FX3D_HD void store_row_down(float* row, uint32_t x, uint32_t W) const { row[x==0u ? W-1u : x-1u] = f2_lo(p[0]); *reinterpret_cast<unsigned long long*>(row+x) = bits64(make_f2(f2_hi(p[0]), f2_lo(p[1]))); row[x+2u] = f2_hi(p[1]); }
	FX3D_HD void store_shift_up(float* q, float* e) const { q[1] = f2_lo(p[0]); *reinterpret_cast<unsigned long long*>(q+2) = bits64(make_f2(f2_hi(p[0]), f2_lo(p[1]))); *e = f2_hi(p[1]); } // q = position of my vector, e = the element right of it
	FX3D_HD void store_shift_down(float* q, float* e) const { *e = f2_lo(p[0]); *reinterpret_cast<unsigned long long*>(q) = bits64(make_f2(f2_hi(p[0]), f2_lo(p[1]))); q[2] = f2_hi(p[1]); } // e = the element left of it
	FX3D_HD void push_back(uint32_t b) { p[0] = make_f2(f2_hi(p[0]), f2_lo(p[1])); p[1] = make_f2(f2_hi(p[1]), __uint_as_float(b)); }  // {e1,e2,e3,b}
	FX3D_HD void push_front(uint32_t b) { p[1] = make_f2(f2_hi(p[0]), f2_lo(p[1])); p[0] = make_f2(__uint_as_float(b), f2_lo(p[0])); } // {b,e0,e1,e2}
	template<int k> FX3D_HD F2 get_pair() const { return p[k]; }
	template<int k> FX3D_HD void set_pair(F2 v) { p[k] = v; }
	template<int k> FX3D_HD void set_lanes(F2 v, bool lo, bool hi) { p[k] = make_f2(lo ? f2_lo(v) : f2_lo(p[k]), hi ? f2_hi(v) : f2_hi(p[k])); }
	template<int k> FX3D_HD void set_masked(F2 v, uint32_t m) { set_lanes<k>(v, (m&0xFFFFu)!=0u, (m>>16)!=0u); } // m: 0xFFFF per lane to overwrite
	static FX3D_HD uint32_t bits(float v) { return __float_as_uint(v); }
	static FX3D_HD float from_bits(uint32_t b) { return __uint_as_float(b); }
#if defined(FX3D_HOST_EMULATION)
	static FX3D_HD void memcpy_bits(F2& d, unsigned long long v) { std::memcpy(&d, &v, 8); }
	static FX3D_HD unsigned long long bits64(const F2& s) { unsigned long long v; std::memcpy(&v, &s, 8); return v; }
#else
	static FX3D_HD void memcpy_bits(F2& d, unsigned long long v) { d.v = v; }
	static FX3D_HD unsigned long long bits64(const F2& s) { return s.v; }
#endif
};
// 16-bit storage: 4 elements in two 32-bit registers; a register is one lane pair
FX3D_HD F2 unpack_half2_raw(uint32_t r) { return make_f2(__half2float(__ushort_as_half((uint16_t)(r&0xFFFFu))), __half2float(__ushort_as_half((uint16_t)(r>>16)))); }
FX3D_HD uint32_t pack_half2_raw(F2 v) {
#if defined(FX3D_HOST_EMULATION)
	return (uint32_t)__half_as_ushort(__float2half_rn(f2_lo(v)))|((uint32_t)__half_as_ushort(__float2half_rn(f2_hi(v)))<<16);
#else
	uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(f2_hi(v)), "f"(f2_lo(v))); return r; // one instruction converts and packs both lanes
#endif
}
template<int ST> FX3D_HD F2 decode_half_pair(uint32_t r) { // two 16-bit elements -> working scale of Codec<ST>
	if constexpr(ST==ST_FP16S) return unpack_half2_raw(r);
	else return vmul_packed(make_f2(__uint_as_float(fp16c_decode_bits(r&0xFFFFu)), __uint_as_float(fp16c_decode_bits(r>>16))), vsplat<F2>(0x1p112f)); // exact scaling
}
template<int ST> FX3D_HD uint32_t encode_half_pair(F2 v) {
	if constexpr(ST==ST_FP16S) return pack_half2_raw(v);
	else { const F2 s = vmul_rz(v, vsplat<F2>(0x1p-112f)); return fp16c_encode_bits(f2_lo(s))|(fp16c_encode_bits(f2_hi(s))<<16); }
}
template<int ST> struct Pack<ST, 4> {
	uint32_t r[2];
	FX3D_HD void load(const uint16_t* q) { const Raw8 t = *reinterpret_cast<const Raw8*>(q); r[0] = t.x; r[1] = t.y; }
	FX3D_HD void store(uint16_t* q) const { Raw8 t; t.x = r[0]; t.y = r[1]; *reinterpret_cast<Raw8*>(q) = t; }
	FX3D_HD void store_tail(uint16_t* q) const { q[1] = (uint16_t)(r[0]>>16); *reinterpret_cast<uint32_t*>(q+2) = r[1]; }
	FX3D_HD void store_head(uint16_t* q) const { *reinterpret_cast<uint32_t*>(q) = r[0]; q[2] = (uint16_t)(r[1]&0xFFFFu); }
	FX3D_HD uint32_t first_bits() const { return r[0]&0xFFFFu; }
	FX3D_HD uint32_t last_bits() const { return r[1]>>16; }
	FX3D_HD void store_seg_up(uint16_t* row, uint32_t x) const { row[x+1u] = (uint16_t)r[0]; *reinterpret_cast<uint32_t*>(row+x+2u) = (r[0]>>16)|(r[1]<<16); row[x+4u] = (uint16_t)(r[1]>>16); }
	FX3D_HD void store_seg_down(uint16_t* row, uint32_t x) const { *(row+x-1) = (uint16_t)r[0]; *reinterpret_cast<uint32_t*>(row+x) = (r[0]>>16)|(r[1]<<16); row[x+2u] = (uint16_t)(r[1]>>16); }
	FX3D_HD void store_row_up(uint16_t* row, uint32_t x, uint32_t W) const { row[x+1u] = (uint16_t)r[0]; *reinterpret_cast<uint32_t*>(row+x+2u) = (r[0]>>16)|(r[1]<<16); row[x+4u==W ? 0u : x+4u] = (uint16_t)(r[1]>>16); }
	FX3D_HD void store_row_down(uint16_t* row, uint32_t x, uint32_t W) const { row[x==0u ? W-1u : x-1u] = (uint16_t)r[0]; *reinterpret_cast<uint32_t*>(row+x) = (r[0]>>16)|(r[1]<<16); row[x+2u] = (uint16_t)(r[1]>>16); }
	FX3D_HD void store_shift_up(uint16_t* q, uint16_t* e) const { q[1] = (uint16_t)r[0]; *reinterpret_cast<uint32_t*>(q+2) = (r[0]>>16)|(r[1]<<16); *e = (uint16_t)(r[1]>>16); }
	FX3D_HD void store_shift_down(uint16_t* q, uint16_t* e) const { *e = (uint16_t)r[0]; *reinterpret_cast<uint32_t*>(q) = (r[0]>>16)|(r[1]<<16); q[2] = (uint16_t)(r[1]>>16); }
	FX3D_HD void push_back(uint32_t b) { r[0] = (r[0]>>16)|(r[1]<<16); r[1] = (r[1]>>16)|(b<<16); }
	FX3D_HD void push_front(uint32_t b) { r[1] = (r[1]<<16)|(r[0]>>16); r[0] = (r[0]<<16)|(b&0xFFFFu); }
	template<int k> FX3D_HD F2 get_pair() const { return decode_half_pair<ST>(r[k]); } // to the working scale of Codec<ST>
	static FX3D_HD uint32_t encode_pair(F2 v) { return encode_half_pair<ST>(v); }
	template<int k> FX3D_HD void set_pair(F2 v) { r[k] = encode_pair(v); }
	template<int k> FX3D_HD void set_lanes(F2 v, bool lo, bool hi) { const uint32_t m = (lo ? 0x0000FFFFu : 0u)|(hi ? 0xFFFF0000u : 0u); r[k] = (encode_pair(v)&m)|(r[k]&~m); }
	template<int k> FX3D_HD void set_masked(F2 v, uint32_t m) { r[k] = (encode_pair(v)&m)|(r[k]&~m); } // one LOP3; m: 0xFFFF per lane to overwrite
	static FX3D_HD uint32_t bits(uint16_t v) { return (uint32_t)v; }
	static FX3D_HD uint16_t from_bits(uint32_t b) { return (uint16_t)b; }
};

// ---- 2 consecutive x-elements of one slot: exactly one F2 lane pair ----
template<> struct Pack<ST_FP32, 2> {
	F2 p;
	FX3D_HD void load(const float* q) { Pack<ST_FP32, 4>::memcpy_bits(p, *reinterpret_cast<const unsigned long long*>(q)); }
	FX3D_HD void store(float* q) const { *reinterpret_cast<unsigned long long*>(q) = Pack<ST_FP32, 4>::bits64(p); }
	FX3D_HD void store_tail(float* q) const { q[1] = f2_hi(p); }
	FX3D_HD void store_head(float* q) const { q[0] = f2_lo(p); }
	FX3D_HD uint32_t first_bits() const { return __float_as_uint(f2_lo(p)); }
	FX3D_HD uint32_t last_bits() const { return __float_as_uint(f2_hi(p)); }
	FX3D_HD void store_shift_up(float* q, float* e) const { q[1] = f2_lo(p); *e = f2_hi(p); }   // q = position of my vector, e = the element right of it
	FX3D_HD void store_shift_down(float* q, float* e) const { *e = f2_lo(p); q[0] = f2_hi(p); } // e = the element left of it
	FX3D_HD void push_back(uint32_t b) { p = make_f2(f2_hi(p), __uint_as_float(b)); }
	FX3D_HD void push_front(uint32_t b) { p = make_f2(__uint_as_float(b), f2_lo(p)); }
	template<int k> FX3D_HD F2 get_pair() const { return p; }
	template<int k> FX3D_HD void set_pair(F2 v) { p = v; }
	template<int k> FX3D_HD void set_lanes(F2 v, bool lo, bool hi) { p = make_f2(lo ? f2_lo(v) : f2_lo(p), hi ? f2_hi(v) : f2_hi(p)); }
	template<int k> FX3D_HD void set_masked(F2 v, uint32_t m) { set_lanes<k>(v, (m&0xFFFFu)!=0u, (m>>16)!=0u); }
	static FX3D_HD uint32_t bits(float v) { return __float_as_uint(v); }
	static FX3D_HD float from_bits(uint32_t b) { return __uint_as_float(b); }
};
template<int ST> struct Pack<ST, 2> {
	uint32_t r;
	FX3D_HD void load(const uint16_t* q) { r = *reinterpret_cast<const uint32_t*>(q); }
	FX3D_HD void store(uint16_t* q) const { *reinterpret_cast<uint32_t*>(q) = r; }
	FX3D_HD void store_tail(uint16_t* q) const { q[1] = (uint16_t)(r>>16); }
	FX3D_HD void store_head(uint16_t* q) const { q[0] = (uint16_t)(r&0xFFFFu); }
	FX3D_HD uint32_t first_bits() const { return r&0xFFFFu; }
	FX3D_HD uint32_t last_bits() const { return r>>16; }
	FX3D_HD void store_shift_up(uint16_t* q, uint16_t* e) const { q[1] = (uint16_t)r; *e = (uint16_t)(r>>16); }
	FX3D_HD void store_shift_down(uint16_t* q, uint16_t* e) const { *e = (uint16_t)r; q[0] = (uint16_t)(r>>16); }
	FX3D_HD void push_back(uint32_t b) { r = (r>>16)|(b<<16); }
	FX3D_HD void push_front(uint32_t b) { r = (r<<16)|(b&0xFFFFu); }
	template<int k> FX3D_HD F2 get_pair() const { return decode_half_pair<ST>(r); }
	template<int k> FX3D_HD void set_pair(F2 v) { r = encode_half_pair<ST>(v); }
	template<int k> FX3D_HD void set_lanes(F2 v, bool lo, bool hi) { const uint32_t m = (lo ? 0x0000FFFFu : 0u)|(hi ? 0xFFFF0000u : 0u); r = (encode_half_pair<ST>(v)&m)|(r&~m); }
	template<int k> FX3D_HD void set_masked(F2 v, uint32_t m) { r = (encode_half_pair<ST>(v)&m)|(r&~m); }
	static FX3D_HD uint32_t bits(uint16_t v) { return (uint32_t)v; }
	static FX3D_HD uint16_t from_bits(uint32_t b) { return (uint16_t)b; }
};

// ================================================================================================================
// stream_collide, vector form: one thread owns K (2 or 4) x-consecutive cells and moves every slot with one aligned
// access of K elements (K=4: 16 bytes FP32 / 8 bytes FP16; K=2: 8 / 4 bytes). Directions with an x component are misaligned by one element: the thread
// loads the aligned vector and obtains / hands over the straddling element by a warp shuffle; only at warp, block,
// row or region ends does a lane fall back to one scalar access. The cell pairs are collided in packed
// binary32x2 arithmetic (F2). Every (cell,slot) is still read and written by exactly one thread (the Esoteric-Pull
// invariant); solid cells' populations pass through unchanged.
// ================================================================================================================
template<int Q, int COLL, int ST, bool VF, int K, int ODD>
__global__ void __launch_bounds__(128, (K==4 ? FX3D_V4_MINBLOCKS : FX3D_V2_MINBLOCKS)) k_stream_collide_vec(const Lattice L, const Region R) {
	constexpr unsigned FULL = 0xFFFFFFFFu;
	typedef Codec<ST> C;
	typedef typename C::elem_t E;
	typedef Pack<ST, K> P;
	const uint32_t g = R.g0+blockIdx.x*blockDim.x+threadIdx.x;
	const uint32_t y = R.y0+blockIdx.y*blockDim.y+threadIdx.y;
	const uint32_t z = R.z0+blockIdx.z;
	const bool valid = g<R.g1 && y<R.y1;
	if(!__any_sync(FULL, valid)) return; // warp-uniform
	const uint32_t lane = (threadIdx.x+threadIdx.y*blockDim.x)&31u;
	const bool has_right = valid && lane<31u && threadIdx.x+1u<blockDim.x && g+1u<R.g1; // lane+1 owns the next K cells of this row
	const bool has_left = valid && lane>0u && threadIdx.x>0u;                           // lane-1 owns the previous K cells
	// Address of my vector in row (y+ey, z+ez) of slot s: one 64-bit pointer per distinct neighbour row, computed once,
	// plus s*slot as a single 32x32+64 multiply-add (IMAD.WIDE.U32) on uniform operands. Lanes outside the region
	// alias the last valid cell group for loads (clamped coordinates) and never store.
	const uint32_t gc = g<R.g1 ? g : R.g1-1u, yc = y<R.y1 ? y : R.y1-1u;
	const uint32_t x0 = L.Hx+(uint32_t)K*gc;
	const uint32_t yy[3] = { dec(yc, L.Ny), yc, inc(yc, L.Ny) }, zz[3] = { dec(z, L.Nz), z, inc(z, L.Nz) };
	const int dxr = (x0+(uint32_t)K>=L.Nx ? 0 : (int)x0+K)-(int)x0; // element offset of the cell right of my vector (periodic wrap)
	const int dxl = (x0==0u ? (int)L.Nx-1 : (int)x0-1)-(int)x0;     // element offset of the cell left of my vector
	constexpr uint32_t odd = (uint32_t)ODD; // step parity t&1 is a template parameter: slot numbers become immediates
	char* rowp[3][3];
	static_for<0, 9, 1>([&](auto J) { constexpr int j = J; rowp[j/3][j%3] = reinterpret_cast<char*>(L.fi)+(row(L, yy[j/3], zz[j%3])+(uint64_t)(x0+L.xo))*sizeof(E); });
	auto at = [&](auto EY, auto EZ, uint32_t s) -> E* { return reinterpret_cast<E*>(mad_wide(L.slot32, s*(uint32_t)sizeof(E), rowp[EY.value+1][EZ.value+1])); };
#define FX3D_AT(ey, ez, s) at(std::integral_constant<int, (ey)>{}, std::integral_constant<int, (ez)>{}, (s))

	uint32_t fl[K];
	bool any_active = false;
	if(valid) {
		const uint8_t* fp = L.flags+lin(L, x0, yc, z);
		if constexpr(K==4) {
			if(((L.Nx|x0)&3u)==0u) { // rows are 4-byte aligned: one load for the 4 flags
				const uint32_t f4 = *reinterpret_cast<const uint32_t*>(fp);
				fl[0] = f4&0xFFu; fl[1] = (f4>>8)&0xFFu; fl[2] = (f4>>16)&0xFFu; fl[3] = f4>>24;
			} else { fl[0] = fp[0]; fl[1] = fp[1]; fl[2] = fp[2]; fl[3] = fp[3]; }
		} else {
			if(((L.Nx|x0)&1u)==0u) { const uint32_t f2 = *reinterpret_cast<const uint16_t*>(fp); fl[0] = f2&0xFFu; fl[1] = f2>>8; }
			else { fl[0] = fp[0]; fl[1] = fp[1]; }
		}
		static_for<0, K, 1>([&](auto J) { any_active = any_active || (fl[J]&TYPE_BO)!=TYPE_S; });
	} else { static_for<0, K, 1>([&](auto J) { fl[J] = TYPE_S; }); }
	// ---- stream in: A[i] holds, per owned cell, the population that load_f() assigns to fhn[i] ----
	P A[Q];
	A[0].load(FX3D_AT(0, 0, 0u));
	static_for<1, Q, 2>([&](auto I) {
		constexpr int i = I;
		A[i  ].load(FX3D_AT(0, 0, odd ? (uint32_t)i : (uint32_t)i+1u));
		A[i+1].load(FX3D_AT(dir_y(i), dir_z(i), odd ? (uint32_t)i+1u : (uint32_t)i)); // aligned vector of the neighbour row; x-shift fixed up below
	});
	static_for<1, Q, 2>([&](auto I) {
		constexpr int i = I;
		if constexpr(dir_x(i)>0) { // need elements x0+1..x0+4: take the right lane's first element
			uint32_t b = __shfl_down_sync(FULL, A[i+1].first_bits(), 1u);
			if(valid && !has_right) b = P::bits(FX3D_AT(dir_y(i), dir_z(i), odd ? (uint32_t)i+1u : (uint32_t)i)[dxr]);
			A[i+1].push_back(b);
		} else if constexpr(dir_x(i)<0) { // need elements x0-1..x0+2: take the left lane's last element
			uint32_t b = __shfl_up_sync(FULL, A[i+1].last_bits(), 1u);
			if(valid && !has_left) b = P::bits(FX3D_AT(dir_y(i), dir_z(i), odd ? (uint32_t)i+1u : (uint32_t)i)[dxl]);
			A[i+1].push_front(b);
		}
	});

	// ---- collide the two cell pairs in packed arithmetic ----
	static_for<0, K/2, 1>([&](auto Pp) {
		constexpr int p = Pp;
		const uint32_t fb_lo = fl[2*p]&TYPE_BO, fb_hi = fl[2*p+1]&TYPE_BO;
		const bool act_lo = valid && fb_lo!=TYPE_S, act_hi = valid && fb_hi!=TYPE_S;
		if(act_lo || act_hi) {
			F2 f[Q];
			static_for<0, Q, 1>([&](auto I) { f[I] = A[I].template get_pair<p>(); });
			const bool e_lo = L.eb!=0u && act_lo && fb_lo==TYPE_E, e_hi = L.eb!=0u && act_hi && fb_hi==TYPE_E;
			const uint64_t n = lin(L, x0+2u*(uint32_t)p, yc, z), N = cells(L);
			F2 rho_e = vsplat<F2>(1.0f), ux_e = vsplat<F2>(0.0f), uy_e = ux_e, uz_e = ux_e;
			if(e_lo || e_hi) {
				const uint64_t nl = e_lo ? n : n+1ull, nh = e_hi ? n+1ull : n; // only TYPE_E lanes are used
				rho_e = make_f2(L.rho[nl], L.rho[nh]); ux_e = make_f2(L.u[nl], L.u[nh]); uy_e = make_f2(L.u[N+nl], L.u[N+nh]); uz_e = make_f2(L.u[2ull*N+nl], L.u[2ull*N+nh]);
			}
			F2 rhon, uxn, uyn, uzn;
			FX3D_COLLIDE<Q, COLL, VF, F2>(f, C::scale, C::inv_scale, e_lo, e_hi, rho_e, ux_e, uy_e, uz_e, L.fx, L.fy, L.fz, L.w, rhon, uxn, uyn, uzn);
			if(L.upd!=0u) {
				if(act_lo && !e_lo) { L.rho[n] = f2_lo(rhon); L.u[n] = f2_lo(uxn); L.u[N+n] = f2_lo(uyn); L.u[2ull*N+n] = f2_lo(uzn); }
				if(act_hi && !e_hi) { L.rho[n+1ull] = f2_hi(rhon); L.u[n+1ull] = f2_hi(uxn); L.u[N+n+1ull] = f2_hi(uyn); L.u[2ull*N+n+1ull] = f2_hi(uzn); }
			}
			// store_f(): fhn[i] goes to the neighbour-side slot, fhn[i+1] to the local slot
			if(act_lo && act_hi) {
				A[0].template set_pair<p>(f[0]);
				static_for<1, Q, 2>([&](auto I) { constexpr int i = I; A[i+1].template set_pair<p>(f[i]); A[i].template set_pair<p>(f[i+1]); });
			} else { // one of the two cells is solid: its populations keep their bits
				A[0].template set_lanes<p>(f[0], act_lo, act_hi);
				static_for<1, Q, 2>([&](auto I) { constexpr int i = I; A[i+1].template set_lanes<p>(f[i], act_lo, act_hi); A[i].template set_lanes<p>(f[i+1], act_lo, act_hi); });
			}
		}
	});

	// ---- stream out (same addresses as stream in) ----
	if(!__any_sync(FULL, any_active)) return; // nothing but solid cells in this warp: nothing changed
	if(valid) A[0].store(FX3D_AT(0, 0, 0u));
	static_for<1, Q, 2>([&](auto I) {
		constexpr int i = I;
		const uint32_t sl = odd ? (uint32_t)i : (uint32_t)i+1u, sn = odd ? (uint32_t)i+1u : (uint32_t)i;
		if(valid) A[i].store(FX3D_AT(0, 0, sl));
		if constexpr(dir_x(i)==0) {
			if(valid) A[i+1].store(FX3D_AT(dir_y(i), dir_z(i), sn));
		} else if constexpr(dir_x(i)>0) { // my values belong to x0+1..x0+4
			const uint32_t last = A[i+1].last_bits();
			const uint32_t up = __shfl_up_sync(FULL, last, 1u);
			if(valid) {
				E* q = FX3D_AT(dir_y(i), dir_z(i), sn);
				if(!has_right) q[dxr] = P::from_bits(last);
				A[i+1].push_front(up); // {left lane's x0, mine x0+1..x0+3}
				if(has_left) A[i+1].store(q); else A[i+1].store_tail(q);
			}
		} else { // my values belong to x0-1..x0+2
			const uint32_t first = A[i+1].first_bits();
			const uint32_t dn = __shfl_down_sync(FULL, first, 1u);
			if(valid) {
				E* q = FX3D_AT(dir_y(i), dir_z(i), sn);
				if(!has_left) q[dxl] = P::from_bits(first);
				A[i+1].push_back(dn); // {mine x0..x0+2, right lane's x0+3}
				if(has_right) A[i+1].store(q); else A[i+1].store_head(q);
			}
		}
	});
#undef FX3D_AT
}

// ================================================================================================================
// stream_collide, pipelined form: the same per-tile work as the vector kernel, but every 128-thread block is persistent,
// walks a strided list of tiles, and fetches the populations of the tile two iterations ahead with cp.async into its
// shared-memory ring (each thread stages only its own vectors, so no block barrier is needed: cp.async.wait_group
// orders a thread's own copies). Loads therefore cost no registers while in flight and the memory system always has
// (stages-1) tiles per warp outstanding, independent of occupancy. Four cells per thread: 8-byte vectors for 16-bit
// storage, 16-byte vectors for FP32. In-place safety is unchanged: the addresses a tile reads are exactly the addresses it alone writes.
// ================================================================================================================
#ifndef FX3D_PIPE_MINBLOCKS
#define FX3D_PIPE_MINBLOCKS 4
#endif
#ifndef FX3D_PIPE_STAGES
#define FX3D_PIPE_STAGES 2
#endif
#ifndef FX3D_PIPE_COLLIDE // collision formulation of the pipelined kernel: 0 collide_cell (all Q populations and equilibria live), 1 collide_cell_fused
#define FX3D_PIPE_COLLIDE -1 // (relax pair by pair), 2 collide_cell_stream (populations unpacked on demand); -1: per storage, as measured on B200
#endif
template<int Q, int ST> FX3D_HDC constexpr int pipe_collide_mode() { return FX3D_PIPE_COLLIDE>=0 ? FX3D_PIPE_COLLIDE : (ST==ST_FP32 && Q==19) ? 1 : 2; }
constexpr int PIPE_STAGES = FX3D_PIPE_STAGES; // ring depth: tiles in flight per thread = stages-1
template<int Q> FX3D_HDC constexpr bool row_used(int ey, int ez) { if(ey==0&&ez==0) return true; for(int i=1; i<Q; i+=2) if(dir_y(i)==ey&&dir_z(i)==ez) return true; return false; } // neighbour rows the odd directions reach
template<int Q> FX3D_HDC constexpr int x_dirs() { int n = 0; for(int i=1; i<Q; i+=2) if(dir_x(i)!=0) n++; return n; }
template<int Q> FX3D_HDC constexpr int x_dir_rank(int i) { int n = 0; for(int k=1; k<i; k+=2) if(dir_x(k)!=0) n++; return n; } // position of odd direction i among the x-shifted ones
// every thread moves 4 cells per tile: 8-byte vectors for 16-bit storage, 16-byte vectors for FP32 (which then affords only a
// 2-deep ring with 2 blocks per SM). D3Q27 FP32 would leave room for a single block that way, so it moves 2 cells per thread.
template<int Q, int ST> FX3D_HDC constexpr int pipe_cells() { return (ST==ST_FP32 && Q>19) ? 2 : 4; }
template<int Q, int ST> FX3D_HDC constexpr int pipe_vector_bytes() { return pipe_cells<Q, ST>()*(ST==ST_FP32 ? 4 : 2); }
template<int ST> FX3D_HDC constexpr int pipe_stages() { return ST==ST_FP32 ? 2 : PIPE_STAGES; }
template<int Q, int ST> FX3D_HDC constexpr int pipe_blocks_per_sm() { return ST==ST_FP32 ? (Q>19 ? 3 : 2) : (Q>19 && FX3D_PIPE_MINBLOCKS>3) ? 3 : FX3D_PIPE_MINBLOCKS; } // also the register cap the kernel is compiled for
template<int Q, int ST> FX3D_HDC constexpr uint32_t pipe_smem_bytes() { return (uint32_t)pipe_stages<ST>()*((uint32_t)Q*128u*(uint32_t)pipe_vector_bytes<Q, ST>()+((uint32_t)x_dirs<Q>()+2u)*128u*4u); } // vectors, edge words, two flag words

FX3D_HD void cp_async8(void* smem_dst, const void* gmem_src) {
#if defined(FX3D_HOST_EMULATION)
	std::memcpy(smem_dst, gmem_src, 8);
#else
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
#endif
}
FX3D_HD void cp_async16(void* smem_dst, const void* gmem_src) {
#if defined(FX3D_HOST_EMULATION)
	std::memcpy(smem_dst, gmem_src, 16);
#else
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
#endif
}
template<int BYTES> FX3D_HD void cp_async_vec(void* smem_dst, const void* gmem_src) { if constexpr(BYTES==16) cp_async16(smem_dst, gmem_src); else cp_async8(smem_dst, gmem_src); }
FX3D_HD void cp_async4(void* smem_dst, const void* gmem_src) {
#if defined(FX3D_HOST_EMULATION)
	std::memcpy(smem_dst, gmem_src, 4);
#else
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
#endif
}
FX3D_HD void cp_async_commit() {
#if !defined(FX3D_HOST_EMULATION)
	asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template<int N> FX3D_HD void cp_async_wait() {
#if !defined(FX3D_HOST_EMULATION)
	asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
#endif
}
FX3D_HD unsigned char* dynamic_smem() {
#if defined(FX3D_HOST_EMULATION)
	return emul::block_smem();
#else
	extern __shared__ __align__(16) unsigned char fx3d_dynamic_smem[];
	return fx3d_dynamic_smem;
#endif
}

// ---- MOVING_BOUNDARIES ----
// linear index of the neighbour of (x,y,z) in direction I
template<int I> FX3D_HD uint64_t neighbour_lin(const Lattice& L, uint32_t x, uint32_t y, uint32_t z) { return lin(L, step<dir_x(I)>(x, L.Nx), step<dir_y(I)>(y, L.Ny), step<dir_z(I)>(z, L.Nz)); }
// is any neighbour a TYPE_S cell with non-zero velocity? (src/kernel.cpp:1381-1385, :1442-1446)
template<int Q> FX3D_HD bool next_to_moving_solid(const Lattice& L, uint32_t x, uint32_t y, uint32_t z) {
	const uint64_t N = cells(L);
	bool r = false;
	static_for<1, Q, 1>([&](auto I) {
		const uint64_t j = neighbour_lin<I.value>(L, x, y, z);
		r = r || ((L.flags[j]&TYPE_BO)==TYPE_S && (L.u[j]!=0.0f || L.u[N+j]!=0.0f || L.u[2ull*N+j]!=0.0f));
	});
	return r;
}
// apply_moving_boundaries, src/kernel.cpp:1104-1113: Dirichlet velocity correction of the streamed-in populations of a TYPE_MS cell;
// f at working scale S (S*fma(w6, cu, f) == fma(w6*S, cu, S*f): power-of-two scaling commutes with the rounding)
template<int Q> FX3D_HD void apply_moving_boundaries(const Lattice& L, uint32_t x, uint32_t y, uint32_t z, float (&f)[Q], const float S = 1.0f) {
	const uint64_t N = cells(L);
	static_for<1, Q, 2>([&](auto I) {
		constexpr int i = I;
		const float w6 = (-6.0f*weight<Q>(i))*S;
		uint64_t j = neighbour_lin<i+1>(L, x, y, z);
		if((L.flags[j]&TYPE_BO)==TYPE_S) f[i  ] = fmaf(w6, (float)dir_x(i+1)*L.u[j]+(float)dir_y(i+1)*L.u[N+j]+(float)dir_z(i+1)*L.u[2ull*N+j], f[i  ]);
		j = neighbour_lin<i>(L, x, y, z);
		if((L.flags[j]&TYPE_BO)==TYPE_S) f[i+1] = fmaf(w6, (float)dir_x(i  )*L.u[j]+(float)dir_y(i  )*L.u[N+j]+(float)dir_z(i  )*L.u[2ull*N+j], f[i+1]);
	});
}
// ---- collision of the K cells a thread holds in A (raw storage vectors, stream-in order), pair by pair in packed arithmetic;
// on return A holds what streams out through the same slots. flags4: the flag bytes of the cells (TYPE_S for cells to skip).
template<int Q, int COLL, int ST, bool VF, int K, bool SG = false, bool MB = false> FX3D_HD void collide_tile(const Lattice& L, Pack<ST, K> (&A)[Q], const uint32_t flags4, const uint32_t x0, const uint32_t yc, const uint32_t z) {
	typedef Codec<ST> C;
	typedef typename C::elem_t E;
	static_for<0, K/2, 1>([&](auto Pp) {
		constexpr int p = Pp;
		const uint32_t fb_lo = (flags4>>(16*p))&TYPE_BO, fb_hi = (flags4>>(16*p+8))&TYPE_BO;
		const bool act_lo = fb_lo!=TYPE_S, act_hi = fb_hi!=TYPE_S;
		if(act_lo || act_hi) {
			const bool e_lo = L.eb!=0u && fb_lo==TYPE_E, e_hi = L.eb!=0u && fb_hi==TYPE_E;
			const uint64_t n = lin(L, x0+2u*(uint32_t)p, yc, z), N = cells(L);
			F2 rho_e = vsplat<F2>(1.0f), ux_e = vsplat<F2>(0.0f), uy_e = ux_e, uz_e = ux_e;
			if(e_lo || e_hi) {
				const uint64_t nl = e_lo ? n : n+1ull, nh = e_hi ? n+1ull : n;
				rho_e = make_f2(L.rho[nl], L.rho[nh]); ux_e = make_f2(L.u[nl], L.u[nh]); uy_e = make_f2(L.u[N+nl], L.u[N+nh]); uz_e = make_f2(L.u[2ull*N+nl], L.u[2ull*N+nh]);
			}
			F2 rhon, uxn, uyn, uzn;
			const bool both = act_lo && act_hi;
			bool moving = false;
			if constexpr(MB) moving = fb_lo==TYPE_MS || fb_hi==TYPE_MS;
			if(moving) { // MOVING_BOUNDARIES, rare lanes: all populations unpacked, the Dirichlet term added lane by lane, then the array collision
				F2 f[Q];
				static_for<0, Q, 1>([&](auto I) { f[I] = A[I].template get_pair<p>(); });
				float fl[Q];
				if(fb_lo==TYPE_MS) {
					static_for<0, Q, 1>([&](auto I) { fl[I] = f2_lo(f[I]); });
					apply_moving_boundaries<Q>(L, x0+2u*(uint32_t)p, yc, z, fl, C::scale);
					static_for<0, Q, 1>([&](auto I) { f[I] = make_f2(fl[I], f2_hi(f[I])); });
				}
				if(fb_hi==TYPE_MS) {
					static_for<0, Q, 1>([&](auto I) { fl[I] = f2_hi(f[I]); });
					apply_moving_boundaries<Q>(L, x0+2u*(uint32_t)p+1u, yc, z, fl, C::scale);
					static_for<0, Q, 1>([&](auto I) { f[I] = make_f2(f2_lo(f[I]), fl[I]); });
				}
				collide_cell<Q, COLL, VF, F2, SG>(f, C::scale, C::inv_scale, e_lo, e_hi, rho_e, ux_e, uy_e, uz_e, L.fx, L.fy, L.fz, L.w, rhon, uxn, uyn, uzn);
				A[0].template set_lanes<p>(f[0], act_lo, act_hi);
				static_for<1, Q, 2>([&](auto I) { constexpr int i = I; A[i+1].template set_lanes<p>(f[i], act_lo, act_hi); A[i].template set_lanes<p>(f[i+1], act_lo, act_hi); });
			} else if constexpr(pipe_collide_mode<Q, ST>()==2) {
			// populations are unpacked on demand and the results packed straight into the slot they stream out through:
			// store_f() sends fhn[i] to the neighbour-side slot (A[i+1]) and fhn[i+1] to the local slot (A[i])
			auto get = [&](auto I) { return A[I.value].template get_pair<p>(); };
			const uint32_t am = (act_lo ? 0x0000FFFFu : 0u)|(act_hi ? 0xFFFF0000u : 0u), em = (e_lo ? 0x0000FFFFu : 0u)|(e_hi ? 0xFFFF0000u : 0u);
			auto put = [&](auto I, F2 v) { // lanes that are not collided here keep what they streamed in
				constexpr int i = I.value;
				constexpr int dst = i==0 ? 0 : (i&1) ? i+1 : i-1;
				if constexpr(sizeof(E)==2) A[dst].template set_masked<p>(v, am);
				else { if(both) A[dst].template set_pair<p>(v); else A[dst].template set_lanes<p>(v, act_lo, act_hi); }
			};
			auto put_e = [&](auto I, F2 v) { // equilibrium-boundary lanes only
				constexpr int i = I.value;
				constexpr int dst = i==0 ? 0 : (i&1) ? i+1 : i-1;
				A[dst].template set_masked<p>(v, em);
			};
			// within a direction pair both members are read before either is written, so the in-place swap is safe
			collide_cell_stream<Q, COLL, VF, F2, SG>(get, put, put_e, C::scale, C::inv_scale, e_lo, e_hi, rho_e, ux_e, uy_e, uz_e, L.fx, L.fy, L.fz, L.w, rhon, uxn, uyn, uzn);
			} else {
			F2 f[Q];
			static_for<0, Q, 1>([&](auto I) { f[I] = A[I].template get_pair<p>(); });
			if constexpr(SG) collide_cell<Q, COLL, VF, F2, true>(f, C::scale, C::inv_scale, e_lo, e_hi, rho_e, ux_e, uy_e, uz_e, L.fx, L.fy, L.fz, L.w, rhon, uxn, uyn, uzn); // SUBGRID: all populations and equilibria at once
			else if constexpr(pipe_collide_mode<Q, ST>()==1) collide_cell_fused<Q, COLL, VF, F2>(f, C::scale, C::inv_scale, e_lo, e_hi, rho_e, ux_e, uy_e, uz_e, L.fx, L.fy, L.fz, L.w, rhon, uxn, uyn, uzn);
			else collide_cell<Q, COLL, VF, F2>(f, C::scale, C::inv_scale, e_lo, e_hi, rho_e, ux_e, uy_e, uz_e, L.fx, L.fy, L.fz, L.w, rhon, uxn, uyn, uzn);
			if(both) {
				A[0].template set_pair<p>(f[0]);
				static_for<1, Q, 2>([&](auto I) { constexpr int i = I; A[i+1].template set_pair<p>(f[i]); A[i].template set_pair<p>(f[i+1]); });
			} else {
				A[0].template set_lanes<p>(f[0], act_lo, act_hi);
				static_for<1, Q, 2>([&](auto I) { constexpr int i = I; A[i+1].template set_lanes<p>(f[i], act_lo, act_hi); A[i].template set_lanes<p>(f[i+1], act_lo, act_hi); });
			}
			}
			if(L.upd!=0u) {
				if(act_lo && !e_lo) { L.rho[n] = f2_lo(rhon); L.u[n] = f2_lo(uxn); L.u[N+n] = f2_lo(uyn); L.u[2ull*N+n] = f2_lo(uzn); }
				if(act_hi && !e_hi) { L.rho[n+1ull] = f2_hi(rhon); L.u[n+1ull] = f2_hi(uxn); L.u[N+n+1ull] = f2_hi(uyn); L.u[2ull*N+n+1ull] = f2_hi(uzn); }
			}
		}
	});
}

template<int Q, int COLL, int ST, bool VF, int ODD>
__global__ void __launch_bounds__(128, pipe_blocks_per_sm<Q, ST>()) k_stream_collide_pipe(const Lattice L, const Region R, const uint32_t tiles_x, const uint32_t tiles_y) {
	// Work unit = a run of z planes of one column of tiles (fixed x-group block and y rows); a block walks its units and, inside
	// a unit, marches in z: the row pointers advance by one plane per tile instead of being rebuilt, and the tile S-1 planes ahead
	// is addressed relative to them (uniform per-plane deltas, which also carry the periodic wrap).
	constexpr int K = pipe_cells<Q, ST>(), VB = pipe_vector_bytes<Q, ST>();
	constexpr int S = pipe_stages<ST>(), NX = x_dirs<Q>();
	constexpr unsigned FULL = 0xFFFFFFFFu;
	constexpr uint32_t odd = (uint32_t)ODD;
	typedef Codec<ST> C;
	typedef typename C::elem_t E;
	typedef Pack<ST, K> P;
	static_assert(sizeof(E)*K==VB, "vector size");
	unsigned char* const smem = dynamic_smem();
	const uint32_t tid = threadIdx.x+threadIdx.y*blockDim.x, lane = tid&31u;
	auto vec_slot = [&](uint32_t stage, int i) -> unsigned char* { return smem+((size_t)(stage*(uint32_t)Q+(uint32_t)i)*128u+tid)*(uint32_t)VB; };
	auto edge_slot = [&](uint32_t stage, int k) -> uint32_t* { return reinterpret_cast<uint32_t*>(smem+(size_t)S*Q*128u*(uint32_t)VB)+(stage*(uint32_t)(NX+2)+(uint32_t)k)*128u+tid; }; // k = NX, NX+1: flag words
	const uint32_t ncols = tiles_x*tiles_y;
	const int64_t plane_bytes = (int64_t)((uint64_t)L.px*L.Ny*sizeof(E)), flag_plane = (int64_t)((uint64_t)L.Nx*L.Ny);
	auto edge_word = [&](const char* q, int d) -> const void* { return reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(q+(int64_t)d*(int64_t)sizeof(E))&~(uintptr_t)3u); };
	auto edge_pick = [&](uint32_t w, const char* q, int d) -> uint32_t { if constexpr(sizeof(E)==4) return w; else return (reinterpret_cast<uintptr_t>(q+(int64_t)d*(int64_t)sizeof(E))&2u) ? w>>16 : w&0xFFFFu; };

	// The (column, plane) tiles, linearised column-major, are cut into gridDim.x equal contiguous shares: every block walks at most
	// one partial column, then whole columns, then one partial column -- perfectly balanced, and only a handful of pipeline fills.
	const uint32_t nz = R.z1-R.z0;
	const uint64_t ntiles = (uint64_t)ncols*nz;
	uint64_t tile = ntiles*blockIdx.x/gridDim.x;
	const uint64_t tile_end = ntiles*(blockIdx.x+1u)/gridDim.x;
	while(tile<tile_end) {
		// ---- column geometry (fixed for the whole unit = the part of one column that belongs to this block) ----
		const uint32_t col = (uint32_t)(tile/nz), zoff = (uint32_t)(tile%nz), xb = col%tiles_x, yb = col/tiles_x;
		const uint32_t zs = R.z0+zoff, ze = (uint64_t)(nz-zoff)<=tile_end-tile ? R.z1 : zs+(uint32_t)(tile_end-tile);
		tile += ze-zs;
		const uint32_t g = R.g0+xb*blockDim.x+threadIdx.x, y = R.y0+yb*blockDim.y+threadIdx.y;
		const bool valid = g<R.g1 && y<R.y1;
		const bool has_right = valid && lane<31u && threadIdx.x+1u<blockDim.x && g+1u<R.g1;
		const bool has_left = valid && lane>0u && threadIdx.x>0u;
		const uint32_t gc = g<R.g1 ? g : R.g1-1u, yc = y<R.y1 ? y : R.y1-1u;
		const uint32_t x0 = L.Hx+(uint32_t)K*gc;
		const uint32_t yy[3] = { dec(yc, L.Ny), yc, inc(yc, L.Ny) };
		const int dxr = (x0+(uint32_t)K>=L.Nx ? 0 : (int)x0+K)-(int)x0, dxl = (x0==0u ? (int)L.Nx-1 : (int)x0-1)-(int)x0;
		const uint8_t* const flag_col = L.flags+((uint64_t)x0+(uint64_t)yc*L.Nx);
		// colp[ey] points at my vector in row (yc+ey) of the current plane z; slot s is added with one IMAD.WIDE, the neighbour
		// planes z-1 / z+1 (periodic wrap included) with a uniform byte offset
		char* colp[3];
		auto locate = [&](uint32_t z) {
			static_for<0, 3, 1>([&](auto J) { colp[J] = reinterpret_cast<char*>(L.fi)+(row(L, yy[J], z)+(uint64_t)(x0+L.xo))*sizeof(E); });
		};
		auto plane_delta = [&](uint32_t zfrom, uint32_t zto) -> int64_t { return ((int64_t)zto-(int64_t)zfrom)*plane_bytes; }; // uniform
#define FX3D_AT(ey, s) mad_wide(L.slot32, (s)*(uint32_t)sizeof(E), colp[(ey)+1])
		auto issue = [&](uint32_t zcur, uint32_t za, uint32_t stage) { // start the copies of the tile at plane za, addressed relative to the pointers of plane zcur
			if(!valid) return;
			const int64_t dz[3] = { plane_delta(zcur, dec(za, L.Nz)), plane_delta(zcur, za), plane_delta(zcur, inc(za, L.Nz)) };
			const uintptr_t fa = reinterpret_cast<uintptr_t>(flag_col+(int64_t)za*flag_plane);
			cp_async4(edge_slot(stage, NX), reinterpret_cast<const void*>(fa&~(uintptr_t)3u));
			if((fa&3u)+(uintptr_t)K>4u) cp_async4(edge_slot(stage, NX+1), reinterpret_cast<const void*>((fa&~(uintptr_t)3u)+4u));
			cp_async_vec<VB>(vec_slot(stage, 0), FX3D_AT(0, 0u)+dz[1]);
			static_for<1, Q, 2>([&](auto I) {
				constexpr int i = I;
				cp_async_vec<VB>(vec_slot(stage, i), FX3D_AT(0, odd ? (uint32_t)i : (uint32_t)i+1u)+dz[1]);
				const char* q = FX3D_AT(dir_y(i), odd ? (uint32_t)i+1u : (uint32_t)i)+dz[dir_z(i)+1];
				cp_async_vec<VB>(vec_slot(stage, i+1), q);
				if constexpr(dir_x(i)>0) { if(!has_right) cp_async4(edge_slot(stage, x_dir_rank<Q>(i)), edge_word(q, dxr)); }
				else if constexpr(dir_x(i)<0) { if(!has_left) cp_async4(edge_slot(stage, x_dir_rank<Q>(i)), edge_word(q, dxl)); }
			});
		};

		locate(zs);
		for(uint32_t k=0u; k<(uint32_t)(S-1); k++) { if(zs+k<ze) issue(zs, zs+k, k); cp_async_commit(); }
		for(uint32_t z=zs, it=0u; z<ze; z++, it++) {
			const uint32_t stage = it%(uint32_t)S;
			if(z+(uint32_t)(S-1)<ze) issue(z, z+(uint32_t)(S-1), (it+(uint32_t)(S-1))%(uint32_t)S);
			const int64_t dzn[3] = { plane_delta(z, dec(z, L.Nz)), 0, plane_delta(z, inc(z, L.Nz)) }; // neighbour planes of this tile
			cp_async_commit();
			cp_async_wait<S-1>(); // the copies of plane z have landed
			uint32_t flags4; // the flag bytes of my 4 cells, kept packed (one register); lanes outside the region carry TYPE_S
			{
				const uint32_t sh = 8u*(uint32_t)(reinterpret_cast<uintptr_t>(flag_col+(int64_t)z*flag_plane)&3u);
				const uint32_t w0 = *edge_slot(stage, NX), w1 = sh+8u*(uint32_t)K>32u ? *edge_slot(stage, NX+1) : 0u;
				flags4 = valid ? (sh==0u ? w0 : (w0>>sh)|(w1<<(32u-sh))) : 0x01010101u*(uint32_t)TYPE_S;
				if constexpr(K==2) flags4 = (flags4&0x0000FFFFu)|(0x01010000u*(uint32_t)TYPE_S); // bytes 2,3 belong to the neighbouring thread
			}
			const bool any_active = ((flags4&(0x01010101u*(uint32_t)TYPE_BO))^(0x01010101u*(uint32_t)TYPE_S))!=0u;

			// ---- stream in from the ring ----
			P A[Q];
			static_for<0, Q, 1>([&](auto I) { A[I].load(reinterpret_cast<const E*>(vec_slot(stage, I))); });
			static_for<1, Q, 2>([&](auto I) {
				constexpr int i = I;
				if constexpr(dir_x(i)>0) {
					uint32_t b = __shfl_down_sync(FULL, A[i+1].first_bits(), 1u);
					if(valid && !has_right) b = edge_pick(*edge_slot(stage, x_dir_rank<Q>(i)), FX3D_AT(dir_y(i), odd ? (uint32_t)i+1u : (uint32_t)i)+dzn[dir_z(i)+1], dxr);
					A[i+1].push_back(b);
				} else if constexpr(dir_x(i)<0) {
					uint32_t b = __shfl_up_sync(FULL, A[i+1].last_bits(), 1u);
					if(valid && !has_left) b = edge_pick(*edge_slot(stage, x_dir_rank<Q>(i)), FX3D_AT(dir_y(i), odd ? (uint32_t)i+1u : (uint32_t)i)+dzn[dir_z(i)+1], dxl);
					A[i+1].push_front(b);
				}
			});

			collide_tile<Q, COLL, ST, VF, K>(L, A, flags4, x0, yc, z);

			// ---- stream out straight from registers (same addresses as stream in) ----
			if(__any_sync(FULL, any_active)) {
				if(valid) A[0].store(reinterpret_cast<E*>(FX3D_AT(0, 0u)));
				static_for<1, Q, 2>([&](auto I) {
					constexpr int i = I;
					const uint32_t sl = odd ? (uint32_t)i : (uint32_t)i+1u, sn = odd ? (uint32_t)i+1u : (uint32_t)i;
					if(valid) A[i].store(reinterpret_cast<E*>(FX3D_AT(0, sl)));
					if constexpr(dir_x(i)==0) {
						if(valid) A[i+1].store(reinterpret_cast<E*>(FX3D_AT(dir_y(i), sn)+dzn[dir_z(i)+1]));
					} else if constexpr(dir_x(i)>0) {
						const uint32_t last = A[i+1].last_bits();
						const uint32_t up = __shfl_up_sync(FULL, last, 1u);
						if(valid) {
							E* q = reinterpret_cast<E*>(FX3D_AT(dir_y(i), sn)+dzn[dir_z(i)+1]);
							if(!has_right) q[dxr] = P::from_bits(last);
							A[i+1].push_front(up);
							if(has_left) A[i+1].store(q); else A[i+1].store_tail(q);
						}
					} else {
						const uint32_t first = A[i+1].first_bits();
						const uint32_t dn = __shfl_down_sync(FULL, first, 1u);
						if(valid) {
							E* q = reinterpret_cast<E*>(FX3D_AT(dir_y(i), sn)+dzn[dir_z(i)+1]);
							if(!has_left) q[dxl] = P::from_bits(first);
							A[i+1].push_back(dn);
							if(has_right) A[i+1].store(q); else A[i+1].store_head(q);
						}
					}
				});
			}
			static_for<0, 3, 1>([&](auto J) { colp[J] += plane_bytes; }); // next plane (z+1 <= ze-1 < Nz, so no wrap here)
		}
		cp_async_wait<0>();
#undef FX3D_AT
	}
}

// ================================================================================================================
// stream_collide, bulk-copy form: persistent blocks march through (row group, plane) tiles; the DDF rows of a tile move between
// HBM and shared memory with TMA bulk copies (cp.async.bulk, completion on an mbarrier) and the results go back the same way.
// Measured on B200 (tools/microbench/ubench3.cu, ubench4.cu): with 8-16 heavy warps per SM, per-thread cp.async loads + STG stores
// saturate at 5.2 TB/s however deep the ring is, bulk copies reach 6.2 TB/s.
//
// Tile = 128/bx whole x-rows of W = 4*bx non-halo cells (bx threads per row, 4 cells per thread). A row of one slot is one
// contiguous segment, so: one copy per (slot, tile row) in, one out. In shared memory the rows of one slot form a SET of a fixed
// size known at compile time, [pad | 128 threads x 4 cells | pad], so that every per-thread access of the hot loop is register +
// immediate offset (the tile shape is a launch parameter; a stride that depended on it cost 60 multiply-adds per thread and tile):
//   * periodic rows (no x halo): the copies fill the rows; the x-shifted directions read and write at wrapped positions
//   * rows with x halos (x-decomposed domains, one-row tiles only): the row of non-halo cells starts on a pitch boundary (see
//     make_lattice) and is copied as it is; only the neighbour-side buffer of a direction with an x component reaches a halo cell,
//     on the side the direction points to, and its copy is [row | pad] (the halo cell x=Nx-1 is the first element of the tail pad)
//     or [pad | row] (x=0 is the last element of the head pad) -- nothing wraps
// so there are no shuffles, no edge accesses and no per-thread global addresses for the DDFs at all.
//
// Ring of S stages per block (S*stage bytes of shared memory; B blocks per SM): tile k of the block's share uses stage k%S. Per
// tile: wait(full[stage]) -> registers <- stage -> block barrier -> refill the stage of the previous tile (its bulk stores have had
// the stream-in to finish reading it) -> collide -> stage <- registers (in place) -> proxy fence + block barrier -> bulk stores.
// Every row buffer is loaded and stored by the same thread, which waits only for the bulk stores it issued itself. The buffers are
// dealt to the four warps; ONE ELECTED LANE of warp w issues buffers w, w+4, .. from an unrolled list in which everything but the
// row position is a compile-time constant (elect.sync: ptxas then emits each UBLKCP once, with uniform-register addresses, instead
// of a loop over the lanes it believes active). Tiles are numbered y-fastest and CLAIMED by the blocks from a device counter (S tiles
// ahead of the tile in hand), so that the resident blocks sweep neighbouring rows together (DRAM page locality across blocks) and no
// launch waits for its slowest SM (FX3D_ROW_DYNAMIC=0 falls back to a fixed round-robin deal).
//
// Fused y/z halo delivery (replaces LBM::communicate_fi for those axes, src/lbm.cpp:1355-1387): under Esoteric-Pull every
// (slot, row) written in step t has exactly one reader in step t+1 -- the tile at the same row for what was written through the
// neighbour side, the tile at row-e_i for what was written locally. The bulk store of a row therefore goes straight to the
// domain that owns the READER's row as a non-halo row (own memory, or the y/z/diagonal neighbour's over NVLink, P.fi[]), at that
// domain's coordinates; no copy is made anywhere else, and no tile of any domain touches that (slot, row) during the same step.
// Loads always come from own memory. x halos are elements inside a row and stay with the exchange kernels.
// ================================================================================================================
#ifndef FX3D_TMA_STAGES
#define FX3D_TMA_STAGES 2
#endif
// direction components for a run-time direction index, from 2-bit fields of a constant (no table in memory)
FX3D_HDC constexpr unsigned long long pack_dirs(int axis) { unsigned long long m = 0ull; for(int i=0; i<27; i++) m |= (unsigned long long)(dir_c(axis, i)+1)<<(2*i); return m; }
FX3D_HD int dir_rt(int axis, uint32_t i) { constexpr unsigned long long mx = pack_dirs(0), my = pack_dirs(1), mz = pack_dirs(2); return (int)(((axis==0 ? mx : axis==1 ? my : mz)>>(2u*i))&3ull)-1; }
template<int Q, int ST> FX3D_HDC constexpr uint32_t tma_stage_bytes() { return (uint32_t)Q*128u*4u*(ST==ST_FP32 ? 4u : 2u)+128u*4u; } // Q row-buffer sets + flag bytes
template<int Q, int ST> FX3D_HDC constexpr int tma_blocks_per_sm() { return ST==ST_FP32 ? 2 : Q>19 ? 3 : 4; } // by shared memory; also the register cap
template<int Q, int ST> FX3D_HDC constexpr uint32_t tma_smem_bytes() { return 128u+(uint32_t)FX3D_TMA_STAGES*tma_stage_bytes<Q, ST>(); }
#if defined(FX3D_HOST_EMULATION)
// emulation: copies happen at issue time; the barrier word counts completed phases (the issuing thread completes the phase itself)
FX3D_HD void mbar_init(uint64_t* b) { __atomic_store_n(b, 0ull, __ATOMIC_SEQ_CST); }
FX3D_HD void mbar_expect_tx(uint64_t*, uint32_t) {}
FX3D_HD void mbar_phase_done_emulated(uint64_t* b) { __atomic_fetch_add(b, 1ull, __ATOMIC_SEQ_CST); }
FX3D_HD void mbar_wait(uint64_t* b, uint32_t parity) { while((__atomic_load_n(b, __ATOMIC_SEQ_CST)&1ull)==(uint64_t)parity) std::this_thread::yield(); }
FX3D_HD void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t*) { std::memcpy(smem_dst, gmem_src, bytes); }
FX3D_HD void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) { std::memcpy(gmem_dst, smem_src, bytes); }
FX3D_HD void bulk_commit() {}
FX3D_HD void bulk_wait_read() {}
FX3D_HD void bulk_wait_all() {}
FX3D_HD void fence_async_smem() {}
#else
FX3D_HD uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
FX3D_HD void mbar_init(uint64_t* b) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(b)) : "memory"); }
FX3D_HD void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(b)), "r"(bytes) : "memory"); }
FX3D_HD void mbar_phase_done_emulated(uint64_t*) {}
FX3D_HD void mbar_wait(uint64_t* b, uint32_t parity) {
	asm volatile("{\n.reg .pred p;\nFX3D_WAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra FX3D_DONE_%=;\nbra FX3D_WAIT_%=;\nFX3D_DONE_%=:\n}" :: "r"(smem_addr(b)), "r"(parity) : "memory");
}
FX3D_HD void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* b) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(b)) : "memory");
}
FX3D_HD void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) { asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gmem_dst), "r"(smem_addr(smem_src)), "r"(bytes) : "memory"); }
FX3D_HD void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
FX3D_HD void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); } // my bulk stores have finished reading shared memory
FX3D_HD void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
FX3D_HD void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); } // my shared-memory writes become visible to the bulk-copy engine
#endif

// one lane of a converged warp (always the same one for the full mask): the form of "lane 0 only" that lets ptxas issue a uniform-datapath
// instruction (UBLKCP) once instead of looping over the lanes it believes may be active
FX3D_HD bool elect_one(uint32_t lane) {
#if defined(FX3D_HOST_EMULATION) || defined(FX3D_ROW_NO_ELECT)
	return lane==0u;
#else
	uint32_t p; asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xFFFFFFFF;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(p)); (void)lane; return p!=0u;
#endif
}
// The lambdas of the whole-row kernel capture the kernel's locals by reference; if the compiler decides not to inline one of them (it did, for the
// D3Q27 FP32 instances inside the full translation unit: 736 bytes of stack, the wind-tunnel line at half speed) every captured local and the
// Lattice parameter are forced into local memory. Hence: always inline.
#define FX3D_LAMBDA __attribute__((always_inline))
// y/z neighbours of a domain for the fused halo delivery: fi[(dy+1)+3*(dz+1)], [4] = the domain itself; unused entries null
struct RowPeers { void* fi[9]; };
// Byte offset of buffer j of a tile from the tile's own address (slot 0, first tile row), for tiles none of whose rows or neighbour rows wrap
// around or belong to another domain -- all but the tiles next to the y/z faces. Worked out on the host per launch (the slot of a buffer depends
// on the step parity); the kernel reads them from constant memory, one 64-bit uniform load per copy.
struct RowOffsets { long long c[27]; };
template<int Q> inline RowOffsets row_offsets(const Lattice& L, uint32_t esz) {
	RowOffsets O;
	for(int j=0; j<27; j++) {
		const int i = j==0 ? 0 : (j&1) ? j : j-1;
		const bool local = j==0 || (j&1);
		const long long slot = j==0 ? 0 : local ? (L.odd ? i : i+1) : (L.odd ? i+1 : i);
		const long long ey = (local || j>=Q) ? 0 : dir_c(1, i), ez = (local || j>=Q) ? 0 : dir_c(2, i);
		O.c[j] = j<Q ? (slot*(long long)L.slot+ey*(long long)L.px+ez*(long long)L.px*(long long)L.Ny)*(long long)esz : 0ll;
	}
	return O;
}
// Resident blocks per SM of the whole-row kernel (also its register cap, 65536/(128*B)) and the ring depth, measured on B200
// (profiles/r02_row_kernel_tuning.txt).
#ifndef FX3D_ROW_BLOCKS_16
#define FX3D_ROW_BLOCKS_16 4 // D3Q19 with 16-bit storage: 128 registers
#endif
#ifndef FX3D_ROW_BLOCKS_32
#define FX3D_ROW_BLOCKS_32 2 // FP32 (both velocity sets): 255 registers
#endif
#ifndef FX3D_ROW_BLOCKS_27
#define FX3D_ROW_BLOCKS_27 3 // D3Q27 with 16-bit storage: 168 registers
#endif
#ifndef FX3D_ROW_STRIDED
#define FX3D_ROW_STRIDED 1 // tile order of the fixed deal (FX3D_ROW_DYNAMIC=0): 1 = y-fastest, dealt round-robin to the blocks; 0 = one contiguous z-fastest share per block
#endif
#ifndef FX3D_ROW_DYNAMIC
#define FX3D_ROW_DYNAMIC 1 // 1: blocks claim their tiles from a device counter instead of taking every gridDim.x-th one
#endif
#ifndef FX3D_ROW_MAX_STAGES
#define FX3D_ROW_MAX_STAGES 16
#endif
template<int Q, int ST> FX3D_HDC constexpr int row_blocks() { return ST==ST_FP32 ? FX3D_ROW_BLOCKS_32 : (Q>19 ? FX3D_ROW_BLOCKS_27 : FX3D_ROW_BLOCKS_16); }
// Cells per thread K (4 or 2) and threads per block T of the whole-row kernel; a tile is K*T cells. Fewer cells per thread = more warps working on
// one tile: the tile is turned around sooner for the same bytes of shared memory.
#ifndef FX3D_ROW_K_16
#define FX3D_ROW_K_16 4
#endif
#ifndef FX3D_ROW_K_32
#define FX3D_ROW_K_32 4
#endif
#ifndef FX3D_ROW_K_27
#define FX3D_ROW_K_27 4
#endif
#ifndef FX3D_ROW_T_16
#define FX3D_ROW_T_16 (512/FX3D_ROW_K_16)
#endif
#ifndef FX3D_ROW_T_32
#define FX3D_ROW_T_32 (512/FX3D_ROW_K_32)
#endif
#ifndef FX3D_ROW_T_27
#define FX3D_ROW_T_27 (512/FX3D_ROW_K_27)
#endif
template<int Q, int ST> FX3D_HDC constexpr int row_cells() { return ST==ST_FP32 ? FX3D_ROW_K_32 : (Q>19 ? FX3D_ROW_K_27 : FX3D_ROW_K_16); }
template<int Q, int ST> FX3D_HDC constexpr uint32_t row_threads() { return ST==ST_FP32 ? FX3D_ROW_T_32 : (Q>19 ? FX3D_ROW_T_27 : FX3D_ROW_T_16); }
constexpr uint32_t ROW_MAX_STAGES = FX3D_ROW_MAX_STAGES<16 ? FX3D_ROW_MAX_STAGES : 16u, ROW_BARRIERS = 128u;
// pad on either side of a row buffer: a whole 32-byte sector, so that a store that ends in a pad ends on a sector boundary -- except for D3Q27 FP32, whose
// two stages of 27 sets fit the shared memory of two blocks per SM only with 16-byte pads
template<int Q, int ST> FX3D_HDC constexpr uint32_t row_pad() { return (Q>19 && ST==ST_FP32) ? 16u : 32u; }
constexpr uint32_t ROW_CLAIMS = 128u; // ring of 32 claimed tile numbers (FX3D_ROW_DYNAMIC)
template<int Q, int ST> FX3D_HDC constexpr uint32_t row_header() { return ROW_BARRIERS+ROW_CLAIMS+2u*row_threads<Q, ST>()*8u; } // header: full[<=16] mbarriers, then two buffers of T x 2 flag words (rows whose flags cannot travel by bulk copy)
template<int Q, int ST> FX3D_HDC constexpr uint32_t row_set_bytes() { return row_threads<Q, ST>()*(uint32_t)row_cells<Q, ST>()*(ST==ST_FP32 ? 4u : 2u)+2u*row_pad<Q, ST>(); } // one slot's rows of a tile: [pad | K*T elements | pad]
template<int Q, int ST> FX3D_HDC constexpr uint32_t row_stage_bytes() { return (uint32_t)Q*row_set_bytes<Q, ST>()+row_threads<Q, ST>()*(uint32_t)row_cells<Q, ST>()+16u; } // Q sets + the flag bytes of the tile (+16: a flag row that starts off a 16-byte boundary)

template<int Q, int COLL, int ST, bool VF, int ODD, bool SG = false, bool MB = false>
__global__ void __launch_bounds__(row_threads<Q, ST>(), row_blocks<Q, ST>()) k_stream_collide_tma(const Lattice L, const Region R, const uint32_t tiles_y, const uint32_t S, const RowPeers P, const RowOffsets O, uint32_t* const tile_counter) {
	constexpr int K = row_cells<Q, ST>();
	constexpr uint32_t T = row_threads<Q, ST>(), NW = T/32u; // threads and warps per block
	static_assert(K==4 && T==128u, "measured on B200: two cells per thread (256 threads per tile) gains nothing for FP32 and loses a third for 16-bit storage; 64-thread blocks for 256-cell rows are no faster either");
	typedef Codec<ST> C;
	typedef typename C::elem_t E;
	typedef Pack<ST, K> PK;
	constexpr uint32_t VB = (uint32_t)(sizeof(E)*K), ESZ = (uint32_t)sizeof(E), PAD = row_pad<Q, ST>(), SET = row_set_bytes<Q, ST>(), STAGE = row_stage_bytes<Q, ST>();
	unsigned char* const smem = dynamic_smem();
	uint64_t* const full = reinterpret_cast<uint64_t*>(smem); // full[stage]: the bulk loads of the tile in this stage have landed
	unsigned char* const ring = smem+row_header<Q, ST>();
	const uint32_t bx = blockDim.x, by = blockDim.y;
	const uint32_t t = threadIdx.x+threadIdx.y*bx, lane = t&31u;
	const uint32_t warp = __shfl_sync(0xFFFFFFFFu, t>>5, 0); // uniform
	const uint32_t W = bx*(uint32_t)K, row_bytes = bx*VB;
	const bool hx = L.Hx!=0u, hy = L.Hy!=0u, hz = L.Hz!=0u; // hx: one-row tiles only (by==1), see the launcher
	// where a row's copy starts in its pitch row (bytes), how much it covers, and where that lands in the set
	// Only the neighbour-side buffer of a direction with an x component reaches a halo cell, and only on the side the direction points to: its copy is
	// [row | pad] (e_x > 0) or [pad | row] (e_x < 0); every other buffer moves the row alone, which starts on a 128-byte line. (Copying pad + row + pad for
	// all buffers, as the first version did, touched 10 lines per copy instead of 8: x-decomposed domains ran 25 % behind periodic ones.)
	const uint32_t g_row = hx ? (L.xo+1u)*ESZ : 0u, x_pad = hx ? PAD : 0u;
	const uint32_t ncopies = (uint32_t)Q*by;
	if(t==0u) { for(uint32_t s=0u; s<S; s++) mbar_init(full+s); }
	fence_async_smem();
	__syncthreads();

	// ---- global address of the copy of buffer j of the tile row at (y, z); a stored row may belong to a y/z neighbour (see the header) ----
	const uint32_t plane = L.px*L.Ny; // elements per z plane of one slot (the padded slot holds < 2^32 elements)
	const bool fused = hy || hz; // some stored rows belong to a y/z neighbour
	auto row_address = [&](bool store, uint32_t slot, bool local, int ey, int ez, uint32_t y, uint32_t z) FX3D_LAMBDA -> char* {
		uint32_t yr = y, zr = z;
		if(!local) { yr = step_rt(ey, y, L.Ny); zr = step_rt(ez, z, L.Nz); }
		char* base = reinterpret_cast<char*>(L.fi);
		if(store && fused) { // (a uniform branch: without y/z halos a store goes where the load came from, and the address stays in uniform registers)
			// the reader's cell row: the row itself for neighbour-side buffers, row-e for local ones
			const uint32_t ry = local ? (uint32_t)((int)y-ey) : yr, rz = local ? (uint32_t)((int)z-ez) : zr;
			int dy = 0, dz = 0;
			if(hy) dy = ry==0u ? -1 : ry==L.Ny-1u ? 1 : 0;
			if(hz) dz = rz==0u ? -1 : rz==L.Nz-1u ? 1 : 0;
			yr = (uint32_t)((int)yr-dy*(int)(L.Ny-2u)); zr = (uint32_t)((int)zr-dz*(int)(L.Nz-2u));
			if(dy!=0 || dz!=0) base = reinterpret_cast<char*>(P.fi[(dy+1)+3*(dz+1)]);
		}
#if defined(FX3D_ROW_ADDR_CXX)
		return base+((uint64_t)slot*L.slot+(uint64_t)(yr*L.px+zr*plane))*ESZ+g_row;
#else
		return mad_wide(yr*L.px+zr*plane, ESZ, mad_wide(L.slot32, slot*ESZ, base))+g_row; // two IMAD.WIDE: the element offset inside a slot fits 32 bits (address of the first non-halo cell of the row)
#endif
	};
	auto copy_rows = [&](auto LOAD, auto STORE_ELSEWHERE, uint32_t y0, uint32_t z, uint32_t stage) FX3D_LAMBDA { // one lane of every warp: its share of the Q buffers, everything but y and z known at compile time
		constexpr bool load = decltype(LOAD)::value;
		unsigned char* const sb = ring+(size_t)stage*STAGE+PAD; // the row inside set 0
		auto one = [&](auto J) FX3D_LAMBDA {
			constexpr int j = J;
			constexpr int i = j==0 ? 0 : (j&1) ? j : j-1; // odd member of the direction pair
			constexpr uint32_t slot = j==0 ? 0u : (j&1) ? (ODD ? (uint32_t)i : (uint32_t)i+1u) : (ODD ? (uint32_t)i+1u : (uint32_t)i);
			constexpr bool local = j==0 || (j&1);
			constexpr int ey = j==0 ? 0 : dir_y(i), ez = j==0 ? 0 : dir_z(i), ex = local ? 0 : dir_x(i);
			const uint32_t before = ex<0 ? x_pad : 0u, bytes = row_bytes+(ex!=0 ? x_pad : 0u);
			_Pragma("unroll 1") for(uint32_t ty=0u; ty<by; ty++) { // (a uniform loop; one pass for one-row tiles)
				char* gp = row_address(decltype(STORE_ELSEWHERE)::value && j!=0, slot, local, ey, ez, y0+ty, z)-before;
				if constexpr(load) bulk_load(sb+(size_t)j*SET+ty*row_bytes-before, gp, bytes, full+stage); else bulk_store(gp, sb+(size_t)j*SET+ty*row_bytes-before, bytes);
			}
		};
		// one-row tiles away from the y/z faces (no row wraps, none is delivered to a neighbour): tile address + a per-buffer constant. (For many-row tiles
		// the same shortcut measured 9 % SLOWER on 256^3 grids than the general addresses -- not understood, left out.)
		auto plain_one_row = [&](auto J) FX3D_LAMBDA {
			constexpr int j = J;
			constexpr int ex = (j==0 || (j&1)) ? 0 : dir_x(j-1);
			const uint32_t before = ex<0 ? x_pad : 0u, bytes = row_bytes+(ex!=0 ? x_pad : 0u);
			char* gp = reinterpret_cast<char*>(L.fi)+((uint64_t)(y0*L.px+z*plane)*ESZ+g_row)+O.c[j]-before;
			if constexpr(load) bulk_load(sb+(size_t)j*SET-before, gp, bytes, full+stage); else bulk_store(gp, sb+(size_t)j*SET-before, bytes);
		};
#if defined(FX3D_ROW_NO_OFFSETS)
		const bool inner_tile = false;
#else
		const bool inner_tile = by==1u && y0>=2u && y0+3u<=L.Ny && z>=2u && z+3u<=L.Nz;
#endif
		if(inner_tile) static_for<0, (int)NW, 1>([&](auto Wc) FX3D_LAMBDA { if(warp==(uint32_t)Wc.value) static_for<Wc.value, Q, (int)NW>(plain_one_row); });
		else static_for<0, (int)NW, 1>([&](auto Wc) FX3D_LAMBDA { if(warp==(uint32_t)Wc.value) static_for<Wc.value, Q, (int)NW>(one); });
	};
	// flag bytes travel with the tile, one bulk copy per tile row into the tail of the stage: rows that start on 16-byte boundaries (no x halo, Nx a
	// multiple of 16) as they are; the flag row of a one-row tile with x halos as the 16-byte-aligned stretch that contains it (threads then read their 4
	// bytes from two words). Other shapes: per thread, see below.
#if defined(FX3D_ROW_NO_FLAGS_BULK)
	const bool flags_bulk = false, flags_skew = false;
#else
	const bool flags_skew = hx && by==1u;                       // bulk copy of an unaligned flag row
	const bool flags_bulk = flags_skew || (!hx && (L.Nx&15u)==0u);
#endif
	auto flag_row = [&](uint32_t y0, uint32_t z) FX3D_LAMBDA -> uint64_t { return ((uint64_t)y0+(uint64_t)z*L.Ny)*L.Nx+L.Hx; }; // index of the first non-halo flag of the tile
	auto load_tile = [&](uint32_t y0, uint32_t z, uint32_t stage) FX3D_LAMBDA { // every thread of the block calls it
		const uint32_t skew = flags_skew ? (uint32_t)(reinterpret_cast<uintptr_t>(L.flags+flag_row(y0, z))&15u) : 0u; // bytes between the 16-byte boundary before the row and the row
		const uint32_t flag_bytes = flags_skew ? ((skew+W+15u)&~15u) : W;
		if(t==0u) mbar_expect_tx(full+stage, ncopies*row_bytes+by*(uint32_t)x_dirs<Q>()*x_pad+(flags_bulk ? by*flag_bytes : 0u));
		if(elect_one(lane)) {
			copy_rows(std::true_type{}, std::false_type{}, y0, z, stage);
			if(flags_bulk && warp==(uint32_t)Q%NW) {
				const uint8_t* gp = L.flags+flag_row(y0, z)-skew;
				unsigned char* sp = ring+(size_t)stage*STAGE+(size_t)Q*SET;
				_Pragma("unroll 1") for(uint32_t ty=0u; ty<by; ty++, gp += L.Nx, sp += W) bulk_load(sp, gp, flag_bytes, full+stage);
			}
		}
#if defined(FX3D_HOST_EMULATION)
		__syncthreads(); // (emulation: copies happen at issue; the phase completes once every thread has made its copies)
		if(t==0u) mbar_phase_done_emulated(full+stage);
#endif
	};
	auto store_tile = [&](uint32_t y0, uint32_t z, uint32_t stage) FX3D_LAMBDA {
		if(elect_one(lane)) {
			if(fused) copy_rows(std::false_type{}, std::true_type{}, y0, z, stage); else copy_rows(std::false_type{}, std::false_type{}, y0, z, stage);
			bulk_commit();
		}
	};

	// ---- this block's share of the (row group, plane) tiles; tile k of the share uses stage k%S ----
	const uint32_t nz = R.z1-R.z0;
	struct Pos { uint32_t yb, zo; }; // tile row group and plane offset
#if FX3D_ROW_DYNAMIC
	// tiles are numbered y-fastest (neighbouring rows in flight together: DRAM page locality across blocks) and CLAIMED from a device counter that the
	// launcher zeroes: S tiles at the start, one more per tile processed, always S tiles ahead of the tile in hand. SMs that run a little faster simply take
	// more tiles; with a fixed deal the slowest SM finished 2-4 % after the average one (profiles/r02_row_fp16s_512.txt). The claimed numbers travel from
	// thread 0 to the block through a ring of 32 words in shared memory, ordered by the two block barriers of every tile.
	const uint32_t ntiles = tiles_y*nz; // (< 2^32-16: checked by the launcher)
	uint32_t* const claims = reinterpret_cast<uint32_t*>(smem+ROW_BARRIERS);
	auto pos_of = [&](uint32_t tile) FX3D_LAMBDA -> Pos { return Pos{ tile%tiles_y, tile/tiles_y }; };
	if(t==0u) { const uint32_t base = atomicAdd(tile_counter, S); for(uint32_t j=0u; j<S; j++) claims[j] = base+j<ntiles ? base+j : ntiles; }
	__syncthreads();
	for(uint32_t j=0u; j<S; j++) { const uint32_t tile = claims[j]; if(tile<ntiles) { const Pos p = pos_of(tile); load_tile(R.y0+p.yb*by, R.z0+p.zo, j); } } // prologue: the first S tiles
	uint32_t tile_now = claims[0];
	Pos cur = pos_of(tile_now<ntiles ? tile_now : 0u);
	const uint32_t n = tile_now<ntiles ? 1u : 0u; // (only "is there a first tile" is known in advance)
#else
#if FX3D_ROW_STRIDED
	// tiles are numbered y-fastest and dealt round-robin to the blocks: at any moment the resident blocks work on ~gridDim.x neighbouring rows,
	// i.e. every one of the 2Q copy streams of the device sweeps one contiguous window of memory (DRAM page locality across blocks)
	const uint64_t ntiles = (uint64_t)tiles_y*nz;
	const uint32_t n = ntiles>blockIdx.x ? (uint32_t)((ntiles-blockIdx.x+gridDim.x-1u)/gridDim.x) : 0u;
	auto advance = [&](Pos p, uint32_t d) FX3D_LAMBDA -> Pos { const uint32_t yb = p.yb+d*gridDim.x; p.zo += yb/tiles_y; p.yb = yb%tiles_y; return p; }; // (d <= 16 stages, a few hundred blocks: no overflow)
	const Pos first = Pos{ blockIdx.x%tiles_y, blockIdx.x/tiles_y };
#else
	const uint64_t ntiles = (uint64_t)tiles_y*nz, T0 = ntiles*blockIdx.x/gridDim.x, T1 = ntiles*(blockIdx.x+1u)/gridDim.x;
	const uint32_t n = (uint32_t)(T1-T0);
	auto advance = [&](Pos p, uint32_t d) FX3D_LAMBDA -> Pos { p.zo += d; while(p.zo>=nz) { p.zo -= nz; p.yb++; } return p; };
	const Pos first = advance(Pos{ (uint32_t)(T0/nz), (uint32_t)(T0%nz) }, 0u);
#endif
	for(uint32_t j=0u; j<S && j<n; j++) { const Pos p = advance(first, j); load_tile(R.y0+p.yb*by, R.z0+p.zo, j); } // prologue: the first S tiles
	Pos cur = first;
#endif // FX3D_ROW_DYNAMIC
	const uint32_t x0 = (uint32_t)K*threadIdx.x;
	// my 4 flag bytes: the two aligned words that hold them travel one tile ahead with per-thread cp.async into a double-buffered corner of
	// shared memory -- no register holds a load in flight (as plain loads the words were spilled on arrival, and the spill store waited for
	// them: 17 % of all stall samples in an ncu capture of that version)
	const uint64_t flag_plane = (uint64_t)L.Nx*L.Ny;
	const uint8_t* const my_flags = L.flags+((uint64_t)(L.Hx+x0)+(uint64_t)(R.y0+threadIdx.y)*L.Nx+(uint64_t)R.z0*flag_plane);
	const uint32_t flag_rows = by*L.Nx; // bytes between consecutive tile row groups
	uint32_t* const flag_words = reinterpret_cast<uint32_t*>(smem+ROW_BARRIERS+ROW_CLAIMS)+2u*t; // [2 buffers][T threads][2 words]
	auto flag_address = [&](Pos p) FX3D_LAMBDA -> uintptr_t { return reinterpret_cast<uintptr_t>(my_flags+(uint64_t)p.yb*flag_rows+(uint64_t)p.zo*flag_plane); };
	auto request_flags = [&](Pos p, uint32_t buffer) FX3D_LAMBDA {
		const uintptr_t a = flag_address(p);
		cp_async4(flag_words+2u*T*buffer, reinterpret_cast<const void*>(a&~(uintptr_t)3u));
		if(a&3u) cp_async4(flag_words+2u*T*buffer+1u, reinterpret_cast<const void*>((a&~(uintptr_t)3u)+4u));
		cp_async_commit();
	};
	auto take_flags = [&](Pos p, uint32_t buffer) FX3D_LAMBDA -> uint32_t {
		cp_async_wait<0>();
		const uint32_t sh = 8u*(uint32_t)(flag_address(p)&3u), w0 = flag_words[2u*T*buffer];
		return sh==0u ? w0 : (w0>>sh)|(flag_words[2u*T*buffer+1u]<<(32u-sh));
	};
	if(n>0u && !flags_bulk) request_flags(cur, 0u);
	// my vector in set 0 of a stage, and the elements right / left of it that the x-shifted directions reach: the next / previous thread's,
	// or -- at the row ends -- the halo element in the pad (x halos) or the other end of the periodic row
	const uint32_t tb = PAD+t*VB;
	const bool last = threadIdx.x+1u==bx, firstt = threadIdx.x==0u;
	const uint32_t up = (!hx && last) ? tb+VB-row_bytes : tb+VB, dn = (!hx && firstt) ? tb+row_bytes-ESZ : tb-ESZ; // byte offsets in the set
	uint32_t stage = 0u, fill = 0u, prev_stage = 0u; // stage = k%S, fill = k/S (how often the stage has been filled before), prev_stage = (k-1)%S
#if FX3D_ROW_DYNAMIC
	for(uint32_t k=0u; tile_now<ntiles; k++) {
#else
	for(uint32_t k=0u; k<n; k++) {
#endif
		const uint32_t y0 = R.y0+cur.yb*by, y = y0+threadIdx.y, z = R.z0+cur.zo;
		mbar_wait(full+stage, fill&1u);
		unsigned char* const sb = ring+(size_t)stage*STAGE;
		uint32_t flags4;
		if(flags_skew) { // my 4 bytes start `skew` bytes into the copied stretch: two aligned words and a funnel shift
			const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(L.flags+flag_row(y0, z))&15u), sh = 8u*(skew&3u);
			const uint32_t* w = reinterpret_cast<const uint32_t*>(sb+(size_t)Q*SET+(skew&~3u)+t*4u);
			flags4 = sh==0u ? w[0] : (w[0]>>sh)|(w[1]<<(32u-sh));
		} else flags4 = flags_bulk ? *reinterpret_cast<const uint32_t*>(sb+(size_t)Q*SET+t*4u) : take_flags(cur, k&1u);
		// ---- stream in from the stage: my vectors, and for the x-shifted directions the element beyond them ----
		PK A[Q];
		static_for<0, Q, 1>([&](auto I) FX3D_LAMBDA { A[I].load(reinterpret_cast<const E*>(sb+(size_t)I.value*SET+tb)); });
		static_for<1, Q, 2>([&](auto I) FX3D_LAMBDA {
			constexpr int i = I;
			if constexpr(dir_x(i)>0) A[i+1].push_back(PK::bits(*reinterpret_cast<const E*>(sb+(size_t)(i+1)*SET+up)));
			else if constexpr(dir_x(i)<0) A[i+1].push_front(PK::bits(*reinterpret_cast<const E*>(sb+(size_t)(i+1)*SET+dn)));
		});
		__syncthreads(); // everybody has read the stage before anybody writes results into it
		// refill the stage of the previous tile now rather than right after its stores: they have had the stream-in above to finish
		// reading it, so the copying thread rarely waits here
#if FX3D_ROW_DYNAMIC
		uint32_t claimed = 0u;
		if(t==0u) claimed = atomicAdd(tile_counter, 1u); // the tile this block will take S tiles from now (its number is published before the second barrier)
		const uint32_t tile_refill = claims[(k+S-1u)&31u], tile_next = claims[(k+1u)&31u];
		if(k>=1u && tile_refill<ntiles) { const Pos p = pos_of(tile_refill); if(elect_one(lane)) bulk_wait_read(); load_tile(R.y0+p.yb*by, R.z0+p.zo, prev_stage); }
		const Pos next = pos_of(tile_next<ntiles ? tile_next : 0u);
		if(tile_next<ntiles && !flags_bulk) request_flags(next, (k+1u)&1u);
#else
		if(k>=1u && k-1u+S<n) { const Pos p = advance(cur, S-1u); if(elect_one(lane)) bulk_wait_read(); load_tile(R.y0+p.yb*by, R.z0+p.zo, prev_stage); } // (stepping the positions without the divisions measured 4 % slower for FP16S: two more live values, more spills)
		const Pos next = advance(cur, 1u);
		if(k+1u<n && !flags_bulk) request_flags(next, (k+1u)&1u);
#endif
		collide_tile<Q, COLL, ST, VF, K, SG, MB>(L, A, flags4, L.Hx+x0, y, z);
		// ---- stream out into the same row buffers ----
		A[0].store(reinterpret_cast<E*>(sb+tb));
		static_for<1, Q, 2>([&](auto I) FX3D_LAMBDA {
			constexpr int i = I;
			A[i].store(reinterpret_cast<E*>(sb+(size_t)i*SET+tb));
			unsigned char* const q = sb+(size_t)(i+1)*SET;
			if constexpr(dir_x(i)==0) A[i+1].store(reinterpret_cast<E*>(q+tb));
			else if constexpr(dir_x(i)>0) A[i+1].store_shift_up(reinterpret_cast<E*>(q+tb), reinterpret_cast<E*>(q+up));
			else A[i+1].store_shift_down(reinterpret_cast<E*>(q+tb), reinterpret_cast<E*>(q+dn));
		});
		fence_async_smem();
#if FX3D_ROW_DYNAMIC
		if(t==0u) claims[(k+S)&31u] = claimed<ntiles ? claimed : ntiles;
#endif
		__syncthreads();
		store_tile(y0, z, stage);
		cur = next;
#if FX3D_ROW_DYNAMIC
		tile_now = tile_next;
#endif
		prev_stage = stage; stage++;
		if(stage==S) { stage = 0u; fill++; }
	}
	if(elect_one(lane)) bulk_wait_all();
}

// ---- row-segment tiles (rows longer than a tile, or shapes the whole-row kernel does not take) ----
template<int Q, int ST> FX3D_HDC constexpr uint32_t tmaseg_set_bytes() { return 128u*4u*(ST==ST_FP32 ? 4u : 2u)+4u*32u; } // up to 4 tile rows, each padded by 2x16 bytes
template<int Q, int ST> FX3D_HDC constexpr uint32_t tmaseg_smem_bytes() { return 128u+(uint32_t)FX3D_TMA_STAGES*(uint32_t)Q*tmaseg_set_bytes<Q, ST>(); }


// ---- hybrid: bulk loads of row segments into padded row buffers, stream-out straight from registers as in k_stream_collide_pipe ----
// tools/microbench/ubench3.cu: it is the *load* half of the LSU path that saturates at low occupancy (bulk loads + STG stores reach
// the same 6.2 TB/s as bulk loads + bulk stores). Per-thread stores settle the ownership of the row ends by themselves (shuffles
// inside a warp, one scalar store at warp / segment ends), so there is no write-back into the stage, no second barrier, no fix-up.
template<int Q, int COLL, int ST, bool VF, int ODD, bool SG = false, bool MB = false>
__global__ void __launch_bounds__(128, tma_blocks_per_sm<Q, ST>()) k_stream_collide_hyb(const Lattice L, const Region R, const uint32_t tiles_x, const uint32_t tiles_y) {
	constexpr int K = 4, S = FX3D_TMA_STAGES, NXD = x_dirs<Q>();
	constexpr uint32_t odd = (uint32_t)ODD;
	typedef Codec<ST> C;
	typedef typename C::elem_t E;
	typedef Pack<ST, K> P;
	constexpr uint32_t VB = (uint32_t)(sizeof(E)*K), PAD = 16u, CH = PAD/(uint32_t)sizeof(E), SET = tmaseg_set_bytes<Q, ST>(), STAGE = (uint32_t)Q*SET;
	unsigned char* const smem = dynamic_smem();
	uint64_t* const full = reinterpret_cast<uint64_t*>(smem);
	unsigned char* const ring = smem+128;
	const uint32_t tid = threadIdx.x+threadIdx.y*blockDim.x;
	const uint32_t W = blockDim.x*(uint32_t)K, row_bytes = blockDim.x*VB, ROWB = row_bytes+2u*PAD; // a row buffer: [pad | W elements | pad]
	const uint32_t nbuf = (uint32_t)Q*blockDim.y;
	if(tid==0u) { for(int s=0; s<S; s++) mbar_init(full+s); }
	fence_async_smem();
	__syncthreads();

	// buffer c = ty*Q+j of the tile at (X0, rows R.y0+yb*by.., plane z): global address of the segment start, x shift of its direction
	auto buffer_row = [&](uint32_t c, uint32_t X0, uint32_t yb, uint32_t z, uint32_t& smem_off, int& ex) -> char* {
		const uint32_t ty = c/(uint32_t)Q, j = c%(uint32_t)Q, y = R.y0+yb*blockDim.y+ty;
		smem_off = j*SET+ty*ROWB;
		uint32_t slot = j, yr = y, zr = z;
		ex = 0;
		if(j>0u) {
			const uint32_t i = (j&1u) ? j : j-1u;
			if(j&1u) slot = odd ? i : i+1u;
			else { slot = odd ? i+1u : i; yr = step_rt(dir_rt(1, i), y, L.Ny); zr = step_rt(dir_rt(2, i), z, L.Nz); ex = dir_rt(0, i); }
		}
		return reinterpret_cast<char*>(L.fi)+((uint64_t)slot*L.slot+row(L, yr, zr)+(uint64_t)(X0+L.xo))*sizeof(E);
	};
	const uint32_t first_copy = (tid>>5)+4u*(tid&31u);
	// One-row tiles (blockDim.y==1, i.e. segments of 512 cells): lane 0 of warp w issues buffers w, w+4, .. from an unrolled list
	// in which everything but X0, y and z is a compile-time constant; the end-chunk elements are stored by threads n*CH+k.
	const bool one_row = blockDim.y==1u;
	const uint32_t warp = __shfl_sync(0xFFFFFFFFu, tid>>5, 0);
		auto row_ptr = [&](auto J, uint32_t X0, uint32_t y, uint32_t z) -> char* { // segment start of buffer J (compile time) in row y of plane z
		constexpr int j = decltype(J)::value;
		constexpr int i = j==0 ? 0 : (j&1) ? j : j-1;
		constexpr uint32_t slot = j==0 ? 0u : (j&1) ? (ODD ? (uint32_t)i : (uint32_t)i+1u) : (ODD ? (uint32_t)i+1u : (uint32_t)i);
		constexpr int ey = (j==0 || (j&1)) ? 0 : dir_y(i), ez = (j==0 || (j&1)) ? 0 : dir_z(i);
		return reinterpret_cast<char*>(L.fi)+((uint64_t)slot*L.slot+row(L, step<ey>(y, L.Ny), step<ez>(z, L.Nz))+(uint64_t)(X0+L.xo))*sizeof(E);
	};
	auto copy_rows = [&](auto LOAD, uint32_t X0, uint32_t yb, uint32_t z, uint32_t stage) {
		constexpr bool load = decltype(LOAD)::value;
		const uint32_t y = R.y0+yb;
		unsigned char* const sb = ring+(size_t)stage*STAGE;
		auto one = [&](auto J) {
			constexpr int j = J;
			constexpr int ex = (j==0 || (j&1)) ? 0 : dir_x(j-1);
			char* g = row_ptr(J, X0, y, z);
			unsigned char* b = sb+(size_t)j*SET;
			if constexpr(load) {
				if constexpr(ex>0) {
					if(X0+W<L.Nx) bulk_load(b+PAD, g, row_bytes+PAD, full+stage);
					else { bulk_load(b+PAD, g, row_bytes, full+stage); bulk_load(b+PAD+row_bytes, g-(size_t)X0*sizeof(E), PAD, full+stage); }
				} else if constexpr(ex<0) {
					if(X0>0u) bulk_load(b, g-PAD, row_bytes+PAD, full+stage);
					else { bulk_load(b+PAD, g, row_bytes, full+stage); bulk_load(b, g+(size_t)(L.Nx-CH)*sizeof(E), PAD, full+stage); }
				} else bulk_load(b+PAD, g, row_bytes, full+stage);
			} else {
				if constexpr(ex>0) bulk_store(g+PAD, b+2u*PAD, row_bytes-PAD); else if constexpr(ex<0) bulk_store(g, b+PAD, row_bytes-PAD); else bulk_store(g, b+PAD, row_bytes);
			}
		};
		if(warp==0u) static_for<0, Q, 4>(one); else if(warp==1u) static_for<1, Q, 4>(one); else if(warp==2u) static_for<2, Q, 4>(one); else static_for<3, Q, 4>(one);
	};
	auto load_tile = [&](uint32_t X0, uint32_t yb, uint32_t z, uint32_t stage) {
		if(tid==0u) mbar_expect_tx(full+stage, blockDim.y*((uint32_t)Q*row_bytes+(uint32_t)NXD*PAD));
		if(one_row) { if((tid&31u)==0u) copy_rows(std::true_type{}, X0, yb, z, stage); }
		else for(uint32_t c=first_copy; c<nbuf; c+=128u) {
			uint32_t off; int ex;
			char* g = buffer_row(c, X0, yb, z, off, ex);
			unsigned char* b = ring+(size_t)stage*STAGE+off;
			if(ex>0) { // segment + the element one past it
				if(X0+W<L.Nx) bulk_load(b+PAD, g, row_bytes+PAD, full+stage);
				else { bulk_load(b+PAD, g, row_bytes, full+stage); bulk_load(b+PAD+row_bytes, g-(size_t)X0*sizeof(E), PAD, full+stage); } // periodic: the row's first chunk
			} else if(ex<0) { // the element before the segment + segment
				if(X0>0u) bulk_load(b, g-PAD, row_bytes+PAD, full+stage);
				else { bulk_load(b+PAD, g, row_bytes, full+stage); bulk_load(b, g+(size_t)(L.Nx-CH)*sizeof(E), PAD, full+stage); } // periodic: the row's last chunk
			} else bulk_load(b+PAD, g, row_bytes, full+stage);
		}
#if defined(FX3D_HOST_EMULATION)
		__syncthreads();
		if(tid==0u) mbar_phase_done_emulated(full+stage);
#endif
	};
	const uint32_t nz = R.z1-R.z0;
	const uint64_t ntiles = (uint64_t)tiles_x*tiles_y*nz;
	uint64_t tile = ntiles*blockIdx.x/gridDim.x;
	const uint64_t tile_end = ntiles*(blockIdx.x+1u)/gridDim.x;
	uint32_t it = 0u;
	const uint32_t x0 = (uint32_t)K*threadIdx.x; // my first cell within the segment
	while(tile<tile_end) {
		const uint32_t col = (uint32_t)(tile/nz), zoff = (uint32_t)(tile%nz), xb = col%tiles_x, yb = col/tiles_x;
		const uint32_t zs = R.z0+zoff, ze = (uint64_t)(nz-zoff)<=tile_end-tile ? R.z1 : zs+(uint32_t)(tile_end-tile);
		tile += ze-zs;
		const uint32_t X0 = L.Hx+(R.g0+xb*blockDim.x)*(uint32_t)K, y = R.y0+yb*blockDim.y+threadIdx.y;
		const uint8_t* const my_flags = L.flags+((uint64_t)(X0+x0)+(uint64_t)y*L.Nx);
		const uint64_t flag_plane = (uint64_t)L.Nx*L.Ny;
		// my 4 flag bytes come from two aligned words; the words of the next plane are requested a whole tile ahead and only
		// combined when they are needed (combining at the load would wait for them on the spot)
		uint32_t fw0, fw1;
		auto request_flags = [&](uint32_t z) {
			const uintptr_t a = reinterpret_cast<uintptr_t>(my_flags+(uint64_t)z*flag_plane);
			fw0 = *reinterpret_cast<const uint32_t*>(a&~(uintptr_t)3u);
			fw1 = (a&3u) ? *reinterpret_cast<const uint32_t*>((a&~(uintptr_t)3u)+4u) : 0u;
		};
		auto combine_flags = [&](uint32_t z) -> uint32_t {
			const uint32_t sh = 8u*(uint32_t)(reinterpret_cast<uintptr_t>(my_flags+(uint64_t)z*flag_plane)&3u);
			return sh==0u ? fw0 : (fw0>>sh)|(fw1<<(32u-sh));
		};
		// stream-out goes straight from registers (as in k_stream_collide_pipe): my vector in rows y-1, y, y+1 of the current plane
		const uint32_t xg = X0+x0, lane = tid&31u;
		const bool has_right = lane<31u && threadIdx.x+1u<blockDim.x, has_left = lane>0u && threadIdx.x>0u;
		const int dxr = (xg+(uint32_t)K>=L.Nx ? 0 : (int)xg+K)-(int)xg, dxl = (xg==0u ? (int)L.Nx-1 : (int)xg-1)-(int)xg;
		const uint32_t yy[3] = { dec(y, L.Ny), y, inc(y, L.Ny) };
		const int64_t plane_bytes = (int64_t)((uint64_t)L.px*L.Ny*sizeof(E));
		char* colp[3];
		static_for<0, 3, 1>([&](auto J) { colp[J] = reinterpret_cast<char*>(L.fi)+(row(L, yy[J], zs)+(uint64_t)(xg+L.xo))*sizeof(E); });
#define FX3D_AT(ey, s) mad_wide(L.slot32, (s)*(uint32_t)sizeof(E), colp[(ey)+1])
		for(uint32_t k=0u; k<(uint32_t)(S-1); k++) if(zs+k<ze) load_tile(X0, yb, zs+k, (it+k)%(uint32_t)S);
		request_flags(zs);
		for(uint32_t z=zs; z<ze; z++, it++) {
			const uint32_t stage = it%(uint32_t)S;
			const uint32_t flags4 = combine_flags(z);
			mbar_wait(full+stage, (it/(uint32_t)S)&1u);
			unsigned char* const sb = ring+(size_t)stage*STAGE+threadIdx.y*ROWB+PAD;
			P A[Q];
			static_for<0, Q, 1>([&](auto I) { A[I].load(reinterpret_cast<const E*>(sb+(size_t)I.value*SET)+x0); });
			static_for<1, Q, 2>([&](auto I) {
				constexpr int i = I;
				const E* rowp = reinterpret_cast<const E*>(sb+(size_t)(i+1)*SET);
				if constexpr(dir_x(i)>0) A[i+1].push_back(P::bits(rowp[x0+(uint32_t)K]));
				else if constexpr(dir_x(i)<0) A[i+1].push_front(P::bits(*(rowp+x0-1)));
			});
			__syncthreads();
			if(z+(uint32_t)(S-1)<ze) load_tile(X0, yb, z+(uint32_t)(S-1), (it+(uint32_t)(S-1))%(uint32_t)S);
			if(z+1u<ze) request_flags(z+1u);
			collide_tile<Q, COLL, ST, VF, K, SG, MB>(L, A, flags4, X0+x0, y, z);
			{
				constexpr unsigned FULL = 0xFFFFFFFFu;
				const int64_t dzn[3] = { ((int64_t)dec(z, L.Nz)-(int64_t)z)*plane_bytes, 0, ((int64_t)inc(z, L.Nz)-(int64_t)z)*plane_bytes };
				A[0].store(reinterpret_cast<E*>(FX3D_AT(0, 0u)));
				static_for<1, Q, 2>([&](auto I) {
					constexpr int i = I;
					const uint32_t sl = odd ? (uint32_t)i : (uint32_t)i+1u, sn = odd ? (uint32_t)i+1u : (uint32_t)i;
					A[i].store(reinterpret_cast<E*>(FX3D_AT(0, sl)));
					E* q = reinterpret_cast<E*>(FX3D_AT(dir_y(i), sn)+dzn[dir_z(i)+1]);
					if constexpr(dir_x(i)==0) A[i+1].store(q);
					else if constexpr(dir_x(i)>0) {
						const uint32_t last = A[i+1].last_bits();
						const uint32_t up = __shfl_up_sync(FULL, last, 1u);
						if(!has_right) q[dxr] = P::from_bits(last);
						A[i+1].push_front(up);
						if(has_left) A[i+1].store(q); else A[i+1].store_tail(q);
					} else {
						const uint32_t first = A[i+1].first_bits();
						const uint32_t dn = __shfl_down_sync(FULL, first, 1u);
						if(!has_left) q[dxl] = P::from_bits(first);
						A[i+1].push_back(dn);
						if(has_right) A[i+1].store(q); else A[i+1].store_head(q);
					}
				});
			}
			static_for<0, 3, 1>([&](auto J) { colp[J] += plane_bytes; });
		}
#undef FX3D_AT
	}
}

// ================================================================================================================
// scalar cell access shared by the general kernels (any Nx): one thread per cell, addresses as load_f/store_f
// ================================================================================================================
template<int Q, int ST> struct CellIO {
	typedef Codec<ST> C;
	typedef typename C::elem_t E;
	uint64_t here, nb[Q]; // physical offsets (without slot) of the cell and of its "+" neighbours (odd i)
	FX3D_HD void locate(const Lattice& L, uint32_t x, uint32_t y, uint32_t z) {
		here = phys(L, x, y, z);
		static_for<1, Q, 2>([&](auto I) {
			constexpr int i = I;
			nb[i] = phys(L, step<dir_x(i)>(x, L.Nx), step<dir_y(i)>(y, L.Ny), step<dir_z(i)>(z, L.Nz));
		});
	}
	FX3D_HD void pull(const Lattice& L, uint32_t odd, float (&f)[Q]) const { // load_f, src/kernel.cpp:1326-1332
		const E* fi = reinterpret_cast<const E*>(L.fi);
		f[0] = C::decode(fi[here]);
		static_for<1, Q, 2>([&](auto I) {
			constexpr int i = I;
			f[i  ] = C::decode(fi[(uint64_t)(odd ? i : i+1)*L.slot+here]);
			f[i+1] = C::decode(fi[(uint64_t)(odd ? i+1 : i)*L.slot+nb[i]]);
		});
	}
	FX3D_HD void push(const Lattice& L, uint32_t odd, const float (&f)[Q]) const { // store_f, src/kernel.cpp:1333-1339
		E* fi = reinterpret_cast<E*>(L.fi);
		fi[here] = C::encode(f[0]);
		static_for<1, Q, 2>([&](auto I) {
			constexpr int i = I;
			fi[(uint64_t)(odd ? i+1 : i)*L.slot+nb[i]] = C::encode(f[i]);
			fi[(uint64_t)(odd ? i : i+1)*L.slot+here] = C::encode(f[i+1]);
		});
	}
};

FX3D_HD bool region_cell(const Region& R, uint32_t& x, uint32_t& y, uint32_t& z) { // g0/g1 are plain x bounds here
	x = R.g0+blockIdx.x*blockDim.x+threadIdx.x; y = R.y0+blockIdx.y*blockDim.y+threadIdx.y; z = R.z0+blockIdx.z;
	return x<R.g1 && y<R.y1;
}

// update_moving_boundaries, src/kernel.cpp:1432-1450
template<int Q>
__global__ void __launch_bounds__(128) k_update_moving_boundaries(const Lattice L, const Region R) {
	uint32_t x, y, z;
	if(!region_cell(R, x, y, z)) return;
	const uint64_t n = lin(L, x, y, z);
	const uint32_t fn = L.flags[n], fb = fn&TYPE_BO;
	if(fb==TYPE_S || fb==TYPE_E || (fn&TYPE_T)) return;
	L.flags[n] = (uint8_t)(next_to_moving_solid<Q>(L, x, y, z) ? fn|TYPE_MS : fn&~TYPE_MS);
}

// general stream_collide (any grid size): one thread per cell, scalar accesses -- the reference's own access pattern
template<int Q, int COLL, int ST, bool VF, bool SG = false>
__global__ void __launch_bounds__(128) k_stream_collide_v1(const Lattice L, const Region R) {
	uint32_t x, y, z;
	if(!region_cell(R, x, y, z)) return;
	const uint64_t n = lin(L, x, y, z), N = cells(L);
	const uint32_t fb = L.flags[n]&TYPE_BO;
	if(fb==TYPE_S) return;
	CellIO<Q, ST> io; io.locate(L, x, y, z);
	float f[Q];
	io.pull(L, L.odd, f);
	if(L.mb!=0u && fb==TYPE_MS) apply_moving_boundaries<Q>(L, x, y, z, f); // :1477-1479
	const bool is_e = L.eb!=0u && fb==TYPE_E;
	float rho_e = 1.0f, ux_e = 0.0f, uy_e = 0.0f, uz_e = 0.0f;
	if(is_e) { rho_e = L.rho[n]; ux_e = L.u[n]; uy_e = L.u[N+n]; uz_e = L.u[2ull*N+n]; }
	float rhon, uxn, uyn, uzn;
	float fxn = L.fx, fyn = L.fy, fzn = L.fz; // :1494
	if constexpr(VF) { if(L.F) { fxn += L.F[n]; fyn += L.F[N+n]; fzn += L.F[2ull*N+n]; } } // FORCE_FIELD, :1497-1503 (without VOLUME_FORCE the force is never used)
	collide_cell<Q, COLL, VF, float, SG>(f, 1.0f, 1.0f, is_e, false, rho_e, ux_e, uy_e, uz_e, fxn, fyn, fzn, L.w, rhon, uxn, uyn, uzn);
	if(L.upd!=0u && !is_e) { L.rho[n] = rhon; L.u[n] = uxn; L.u[N+n] = uyn; L.u[2ull*N+n] = uzn; }
	io.push(L, L.odd, f);
}

// ---- stream_collide, occupancy form: one cell per thread like the general kernel, but built to stay within 64 registers
// (32 warps per SM) -- 32-bit element offsets formed from three x, three row and three plane terms at the point of use, slot
// numbers as immediates (step parity is a template parameter), the pairwise-fused collision. The reference's own OpenCL kernel
// has this shape (56 registers, 36 warps per SM on the B200's driver compiler) and reaches the full copy bandwidth in FP32, where
// the low-occupancy bulk-copy pipeline tops out at about 6.2 TB/s; for 16-bit storage the bulk-copy kernels are far ahead.
#ifndef FX3D_OCC_MINBLOCKS
#define FX3D_OCC_MINBLOCKS 8
#endif
#ifndef FX3D_OCC_THREADS
#define FX3D_OCC_THREADS 128
#endif
template<int Q, int COLL, int ST, bool VF, int ODD>
__global__ void __launch_bounds__(FX3D_OCC_THREADS, FX3D_OCC_MINBLOCKS) k_stream_collide_occ(const Lattice L, const Region R) {
	typedef Codec<ST> C;
	typedef typename C::elem_t E;
	uint32_t x, y, z;
	if(!region_cell(R, x, y, z)) return;
	const uint32_t n = x+(y+z*L.Ny)*L.Nx; // the reference's uxx=uint index: N < 2^32 (checked on the host)
	const uint32_t fb = L.flags[n]&TYPE_BO;
	if(fb==TYPE_S) return;
	// element offsets inside one slot: ox[ex+1]+oy[ey+1]+oz[ez+1], each < 2^32 in sum (the padded slot holds < 2^32 elements)
	const uint32_t ox[3] = { dec(x, L.Nx)+L.xo, x+L.xo, inc(x, L.Nx)+L.xo };
	const uint32_t oy[3] = { dec(y, L.Ny)*L.px, y*L.px, inc(y, L.Ny)*L.px };
	const uint32_t pz = L.px*L.Ny;
	const uint32_t oz[3] = { dec(z, L.Nz)*pz, z*pz, inc(z, L.Nz)*pz };
	char* const base = reinterpret_cast<char*>(L.fi);
	auto at = [&](auto EX, auto EY, auto EZ, uint32_t slot) -> E* { // one IADD3 + one IMAD.WIDE pair
		const uint32_t off = ox[EX.value+1]+oy[EY.value+1]+oz[EZ.value+1];
		return reinterpret_cast<E*>(mad_wide(off, (uint32_t)sizeof(E), mad_wide(L.slot32, slot*(uint32_t)sizeof(E), base)));
	};
	typedef std::integral_constant<int, 0> Z0;
	float f[Q];
	f[0] = C::decode(load_global<E>(at(Z0{}, Z0{}, Z0{}, 0u)));
	static_for<1, Q, 2>([&](auto I) { // load_f, src/kernel.cpp:1326-1332
		constexpr int i = I;
		f[i  ] = C::decode(load_global<E>(at(Z0{}, Z0{}, Z0{}, ODD ? (uint32_t)i : (uint32_t)i+1u)));
		f[i+1] = C::decode(load_global<E>(at(std::integral_constant<int, dir_x(i)>{}, std::integral_constant<int, dir_y(i)>{}, std::integral_constant<int, dir_z(i)>{}, ODD ? (uint32_t)i+1u : (uint32_t)i)));
	});
	const bool is_e = L.eb!=0u && fb==TYPE_E;
	float rho_e = 1.0f, ux_e = 0.0f, uy_e = 0.0f, uz_e = 0.0f;
	const uint64_t N = cells(L);
	if(is_e) { rho_e = L.rho[n]; ux_e = L.u[n]; uy_e = L.u[N+n]; uz_e = L.u[2ull*N+n]; }
	float rhon, uxn, uyn, uzn;
	collide_cell_fused<Q, COLL, VF, float>(f, 1.0f, 1.0f, is_e, false, rho_e, ux_e, uy_e, uz_e, L.fx, L.fy, L.fz, L.w, rhon, uxn, uyn, uzn);
	if(L.upd!=0u && !is_e) { L.rho[n] = rhon; L.u[n] = uxn; L.u[N+n] = uyn; L.u[2ull*N+n] = uzn; }
	auto at_again = [&](auto EX, auto EY, auto EZ, uint32_t slot) -> E* {
		const uint32_t off = ox[EX.value+1]+oy[EY.value+1]+oz[EZ.value+1];
		return reinterpret_cast<E*>(mad_wide_again(off, (uint32_t)sizeof(E), mad_wide_again(L.slot32, slot*(uint32_t)sizeof(E), base)));
	};
	store_global<E>(at_again(Z0{}, Z0{}, Z0{}, 0u), C::encode(f[0]));
	static_for<1, Q, 2>([&](auto I) { // store_f, src/kernel.cpp:1333-1339
		constexpr int i = I;
		store_global<E>(at_again(std::integral_constant<int, dir_x(i)>{}, std::integral_constant<int, dir_y(i)>{}, std::integral_constant<int, dir_z(i)>{}, ODD ? (uint32_t)i+1u : (uint32_t)i), C::encode(f[i]));
		store_global<E>(at_again(Z0{}, Z0{}, Z0{}, ODD ? (uint32_t)i : (uint32_t)i+1u), C::encode(f[i+1]));
	});
}

// initialize, src/kernel.cpp:1358-1430 (build without MOVING_BOUNDARIES / SURFACE / TEMPERATURE)
template<int Q, int ST>
__global__ void __launch_bounds__(128) k_initialize(const Lattice L, const Region R) {
	uint32_t x, y, z;
	if(!region_cell(R, x, y, z)) return;
	const uint64_t n = lin(L, x, y, z), N = cells(L);
	const uint32_t fn = L.flags[n], fb = fn&TYPE_BO;
	if(L.mb==0u) { if(fb==TYPE_S) { L.u[n] = 0.0f; L.u[N+n] = 0.0f; L.u[2ull*N+n] = 0.0f; } }
	else if(fb==TYPE_S) { // MOVING_BOUNDARIES (:1374-1386): only solids enclosed by solids lose their velocity ...
		bool only_s = true;
		static_for<1, Q, 1>([&](auto I) { only_s = only_s && (L.flags[neighbour_lin<I.value>(L, x, y, z)]&TYPE_BO)==TYPE_S; });
		if(only_s) { L.u[n] = 0.0f; L.u[N+n] = 0.0f; L.u[2ull*N+n] = 0.0f; }
	} else if(fb!=TYPE_E) L.flags[n] = (uint8_t)(next_to_moving_solid<Q>(L, x, y, z) ? fn|TYPE_MS : fn&~TYPE_MS); // ... and cells next to a moving solid are marked
	float feq[Q];
	equilibrium<Q, float>(L.rho[n], L.u[n], L.u[N+n], L.u[2ull*N+n], 1.0f, feq);
	CellIO<Q, ST> io; io.locate(L, x, y, z);
	io.push(L, 1u, feq); // odd-step layout is baked in (:1429)
}

// update_fields, src/kernel.cpp:1794-1870
template<int Q, int ST, bool VF>
__global__ void __launch_bounds__(128) k_update_fields(const Lattice L, const Region R) {
	uint32_t x, y, z;
	if(!region_cell(R, x, y, z)) return;
	const uint64_t n = lin(L, x, y, z), N = cells(L);
	const uint32_t fb = L.flags[n]&TYPE_BO;
	if(fb==TYPE_S) return;
	CellIO<Q, ST> io; io.locate(L, x, y, z);
	float f[Q];
	io.pull(L, L.odd, f);
	if(L.mb!=0u && fb==TYPE_MS) apply_moving_boundaries<Q>(L, x, y, z, f); // :1813-1815
	float rhon, uxn, uyn, uzn;
	float fxn = L.fx, fyn = L.fy, fzn = L.fz;
	if constexpr(VF) { if(L.F) { fxn += L.F[n]; fyn += L.F[N+n]; fzn += L.F[2ull*N+n]; } } // FORCE_FIELD, :1821-1827
	fields_of_cell<Q, VF>(f, fxn, fyn, fzn, rhon, uxn, uyn, uzn);
	if(L.eb!=0u && fb==TYPE_E) return;
	L.rho[n] = rhon; L.u[n] = uxn; L.u[N+n] = uyn; L.u[2ull*N+n] = uzn;
}

// ================================================================================================================
// voxelize_mesh / unvoxelize_mesh, src/kernel.cpp:2267-2357: one thread per cell of the face normal to
// `direction` casts a ray through its column, intersects it with every triangle (bidirectional Moeller-Trumbore, the reference's
// operation order: the flags must come out identical), sorts the hit distances and walks the column toggling inside/outside.
// The triangles are staged through shared memory a chunk at a time -- every thread of the block reads the same triangle at the
// same moment -- together with the two edge vectors p1-p0 and p2-p0, which do not depend on the thread.
// ================================================================================================================
struct Float3 { float x, y, z; };
FX3D_HD Float3 f3(float x, float y, float z) { return Float3{ x, y, z }; }
FX3D_HD Float3 operator+(Float3 a, Float3 b) { return Float3{ a.x+b.x, a.y+b.y, a.z+b.z }; }
FX3D_HD Float3 operator-(Float3 a, Float3 b) { return Float3{ a.x-b.x, a.y-b.y, a.z-b.z }; }
FX3D_HD Float3 cross3(Float3 a, Float3 b) { return Float3{ a.y*b.z-a.z*b.y, a.z*b.x-a.x*b.z, a.x*b.y-a.y*b.x }; }
FX3D_HD float dot3(Float3 a, Float3 b) { return a.x*b.x+a.y*b.y+a.z*b.z; }
FX3D_HD Float3 cell_position(const Lattice& L, uint32_t x, uint32_t y, uint32_t z) { return f3((float)x+0.5f-0.5f*(float)L.Nx, (float)y+0.5f-0.5f*(float)L.Ny, (float)z+0.5f-0.5f*(float)L.Nz); } // :828-830
FX3D_HD int clamp_int(int v, int lo, int hi) { return v<lo ? lo : v>hi ? hi : v; }
struct VoxelizeArgs { // the 16-float block of LBM_Domain::voxelize_mesh_on_device (src/lbm.cpp:279-296) and the domain offset (def_Ox.., src/lbm.cpp:335)
	uint32_t triangle_number; float x0, y0, z0, x1, y1, z1, cx, cy, cz, ux, uy, uz, rx, ry, rz; int Ox, Oy, Oz;
};
constexpr uint32_t VOX_CHUNK = 128u; // triangles staged per pass: 128 x 9 floats x 4 B x (vertices + edges) = 9 KB
template<int Q, int ST>
__global__ void __launch_bounds__(128) k_voxelize_mesh(const Lattice L, const uint32_t direction, const uint32_t t_odd, const uint8_t flag, const float* p0, const float* p1, const float* p2, const VoxelizeArgs V) {
	typedef Codec<ST> C;
	const uint32_t a = blockIdx.x*blockDim.x+threadIdx.x, A = direction==0u ? L.Ny*L.Nz : direction==1u ? L.Nz*L.Nx : L.Nx*L.Ny;
	float* const tri = reinterpret_cast<float*>(dynamic_smem()); // [VOX_CHUNK][p0 | p1-p0 | p2-p0]
	const uint32_t hmin = direction==0u ? (uint32_t)clamp_int((int)V.x0-V.Ox, 0, (int)L.Nx-1) : direction==1u ? (uint32_t)clamp_int((int)V.y0-V.Oy, 0, (int)L.Ny-1) : (uint32_t)clamp_int((int)V.z0-V.Oz, 0, (int)L.Nz-1);
	const uint32_t hmax = direction==0u ? (uint32_t)clamp_int((int)V.x1-V.Ox, 0, (int)L.Nx-1) : direction==1u ? (uint32_t)clamp_int((int)V.y1-V.Oy, 0, (int)L.Ny-1) : (uint32_t)clamp_int((int)V.z1-V.Oz, 0, (int)L.Nz-1);
	const uint32_t ac = a<A ? a : A-1u; // threads past the face keep the block's staging loop company
	uint32_t X, Y, Z;
	if(direction==0u) { X = hmin; Y = ac%L.Ny; Z = ac/L.Ny; } else if(direction==1u) { X = ac/L.Nz; Y = hmin; Z = ac%L.Nz; } else { X = ac%L.Nx; Y = ac/L.Nx; Z = hmin; }
	const Float3 offset = f3(0.5f*(float)((int)L.Nx+2*V.Ox)-0.5f, 0.5f*(float)((int)L.Ny+2*V.Oy)-0.5f, 0.5f*(float)((int)L.Nz+2*V.Oz)-0.5f);
	const Float3 r_origin = cell_position(L, X, Y, Z)+offset;
	const Float3 r_direction = f3((float)(direction==0u), (float)(direction==1u), (float)(direction==2u));
	const bool outside_box = direction==0u ? (r_origin.y<V.y0||r_origin.z<V.z0||r_origin.y>=V.y1||r_origin.z>=V.z1) : direction==1u ? (r_origin.x<V.x0||r_origin.z<V.z0||r_origin.x>=V.x1||r_origin.z>=V.z1) : (r_origin.x<V.x0||r_origin.y<V.y0||r_origin.x>=V.x1||r_origin.y>=V.y1);
	const bool active = a<A && !outside_box;
	uint32_t intersections = 0u, intersections_check = 0u;
	uint16_t distances[64]; // up to 64 mesh intersections per column, as in the reference
	for(uint32_t base=0u; base<V.triangle_number; base+=VOX_CHUNK) {
		const uint32_t count = V.triangle_number-base<VOX_CHUNK ? V.triangle_number-base : VOX_CHUNK;
		__syncthreads();
		for(uint32_t k=threadIdx.x; k<count; k+=blockDim.x) {
			const uint32_t i = 3u*(base+k);
			const Float3 a0 = f3(p0[i], p0[i+1u], p0[i+2u]), eu = f3(p1[i], p1[i+1u], p1[i+2u])-a0, ev = f3(p2[i], p2[i+1u], p2[i+2u])-a0;
			float* q = tri+9u*k;
			q[0] = a0.x; q[1] = a0.y; q[2] = a0.z; q[3] = eu.x; q[4] = eu.y; q[5] = eu.z; q[6] = ev.x; q[7] = ev.y; q[8] = ev.z;
		}
		__syncthreads();
		if(active) for(uint32_t k=0u; k<count; k++) {
			const float* q = tri+9u*k;
			const Float3 p0i = f3(q[0], q[1], q[2]), eu = f3(q[3], q[4], q[5]), ev = f3(q[6], q[7], q[8]);
			const Float3 w = r_origin-p0i, h = cross3(r_direction, ev), qq = cross3(w, eu);
			const float g = dot3(eu, h), f = 1.0f/g, s = f*dot3(w, h), tt = f*dot3(r_direction, qq), d = f*dot3(ev, qq);
			if(g!=0.0f&&s>=0.0f&&s<1.0f&&tt>=0.0f&&s+tt<1.0f) {
				if(d>0.0f) { if(intersections<64u&&d<65536.0f) distances[intersections] = (uint16_t)d; intersections++; }
				else intersections_check++;
			}
		}
	}
	if(!active) return;
	for(uint32_t i=1u; i<(intersections<64u ? intersections : 64u); i++) { // insertion sort
		const uint16_t tv = distances[i];
		uint32_t j = i;
		while(j>0u&&distances[j-1u]>tv) { distances[j] = distances[j-1u]; j--; }
		distances[j] = tv;
	}
	bool inside = (intersections%2u)&&(intersections_check%2u);
	const bool set_u = V.ux*V.ux+V.uy*V.uy+V.uz*V.uz+V.rx*V.rx+V.ry*V.ry+V.rz*V.rz>0.0f;
	uint32_t intersection = intersections%2u!=intersections_check%2u;
	const uint32_t h0 = direction==0u ? X : direction==1u ? Y : Z;
	const uint32_t hmesh = intersections>0u ? h0+(uint32_t)distances[intersections-1u<63u ? intersections-1u : 63u] : 0u; // unused without intersections: inside stays false
	const uint64_t N = cells(L);
	for(uint32_t hh=h0; hh<=hmax; hh++) {
		while(intersection<intersections&&hh>h0+(uint32_t)distances[intersection<63u ? intersection : 63u]) { inside = !inside; intersection++; }
		inside = inside&&(intersection<intersections&&hh<hmesh);
		const uint32_t cx_ = direction==0u ? hh : X, cy_ = direction==1u ? hh : Y, cz_ = direction==2u ? hh : Z;
		const uint64_t n = lin(L, cx_, cy_, cz_);
		uint32_t flagsn = L.flags[n];
		const Float3 p = cell_position(L, cx_, cy_, cz_)+offset;
		const Float3 u_set = f3(V.ux, V.uy, V.uz)+cross3(f3(V.cx, V.cy, V.cz)-p, f3(V.rx, V.ry, V.rz));
		if(inside) {
			flagsn = (flagsn&~TYPE_BO)|flag;
			if(set_u) { L.u[n] = u_set.x; L.u[N+n] = u_set.y; L.u[2ull*N+n] = u_set.z; }
		} else if((flagsn&TYPE_BO)==TYPE_S&&(flagsn&0xC0u)==((uint32_t)flag&0xC0u)) { // the cell was solid, with the same TYPE_X/TYPE_Y marker
			const float unx = L.u[n], uny = L.u[N+n], unz = L.u[2ull*N+n];
			if(unx==u_set.x&&uny==u_set.y&&unz==u_set.z) { // it belonged to this geometry
				if(set_u) { // a moving solid left the cell: its DDFs restart from equilibrium at rho=1 (no mass drift)
					float feq[Q];
					equilibrium<Q, float>(1.0f, unx, uny, unz, 1.0f, feq);
					CellIO<Q, ST> io; io.locate(L, cx_, cy_, cz_);
					io.push(L, t_odd, feq);
				}
				flagsn = (flagsn&TYPE_BO)==TYPE_MS ? flagsn&~TYPE_MS : flagsn&~(uint32_t)flag;
			}
		}
		L.flags[n] = (uint8_t)flagsn;
	}
}
#if defined(FX3D_TU_LBM)
__global__ void __launch_bounds__(128) k_unvoxelize_mesh(const Lattice L, const uint8_t flag, const float x0, const float y0, const float z0, const float x1, const float y1, const float z1, const int Ox, const int Oy, const int Oz) { // :2351-2357
	const uint64_t n = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x;
	if(n>=cells(L)) return;
	const uint64_t plane = (uint64_t)L.Nx*L.Ny, r = n%plane;
	const Float3 p = cell_position(L, (uint32_t)(r%L.Nx), (uint32_t)(r/L.Nx), (uint32_t)(n/plane))+f3(0.5f*(float)((int)L.Nx+2*Ox)-0.5f, 0.5f*(float)((int)L.Ny+2*Oy)-0.5f, 0.5f*(float)((int)L.Nz+2*Oz)-0.5f);
	if(p.x>=x0-1.0f&&p.y>=y0-1.0f&&p.z>=z0-1.0f&&p.x<=x1+1.0f&&p.y<=y1+1.0f&&p.z<=z1+1.0f) L.flags[n] &= (uint8_t)~flag;
}
#endif

FX3D_HD uint32_t face_area_of(const Lattice& L, uint32_t axis) { return axis==0u ? L.Ny*L.Nz : axis==1u ? L.Nz*L.Nx : L.Nx*L.Ny; } // get_area, src/kernel.cpp:2049-2052
FX3D_HD void face_cell_of(const Lattice& L, uint32_t axis, uint32_t a, uint32_t layer, uint32_t& x, uint32_t& y, uint32_t& z) { // :2053-2068
	if(axis==0u) { x = layer; y = a%L.Ny; z = a/L.Ny; } else if(axis==1u) { x = a/L.Nz; y = layer; z = a%L.Nz; } else { x = a%L.Nx; y = a/L.Nx; z = layer; }
}
// ================================================================================================================
// FORCE_FIELD, src/kernel.cpp:1873-1959; host side src/lbm.cpp:206-239,986-1016
// ================================================================================================================
// update_force_field, :1873-1884: the force of the fluid on a solid cell is twice the momentum of the populations streaming into it
// (they bounce back): calculate_rho_u on them, F = 2*Fb*(fx,fy,fz)
template<int Q, int ST>
__global__ void __launch_bounds__(128) k_update_force_field(const Lattice L, const Region R) {
	uint32_t x, y, z;
	if(!region_cell(R, x, y, z)) return;
	const uint64_t n = lin(L, x, y, z), N = cells(L);
	if((L.flags[n]&TYPE_BO)!=TYPE_S) return;
	CellIO<Q, ST> io; io.locate(L, x, y, z);
	float f[Q];
	io.pull(L, L.odd, f);
	float Fb, fx, fy, fz;
	moments<Q, float>(f, 1.0f, 1.0f, Fb, fx, fy, fz);
	const float s = 2.0f*Fb;
	L.F[n] = s*fx; L.F[N+n] = s*fy; L.F[2ull*N+n] = s*fz;
}
// object_center_of_mass (KIND 0) / object_force (1) / object_torque (2), :1901-1959. The reference adds one partial sum per work-group into
// object_sum with floating-point atomics, i.e. in no defined order; here the order is fixed so that results are reproducible (and equal the
// oracle's bit for bit): pass 1 reduces every group of 64 consecutive cells with the reference's stride-doubling tree into partial[group],
// pass 2 (one warp) adds the non-zero partial sums in ascending group order -- one of the orders the reference's atomics can produce.
constexpr uint32_t OBJECT_GROUP = 64u; // the reference's work-group size (src/opencl.hpp WORKGROUP_SIZE)
struct alignas(16) ObjectPartial { float x, y, z; uint32_t cells; };
template<int KIND>
__global__ void __launch_bounds__(OBJECT_GROUP) k_object_partial(const Lattice L, const uint8_t flag_marker, const float cx, const float cy, const float cz, ObjectPartial* partial) {
	float (*cache)[OBJECT_GROUP] = reinterpret_cast<float (*)[OBJECT_GROUP]>(dynamic_smem()); // [3][OBJECT_GROUP] floats, then the cell counts
	uint32_t* count = reinterpret_cast<uint32_t*>(dynamic_smem())+3u*OBJECT_GROUP;
	const uint32_t lid = threadIdx.x;
	const uint64_t n = (uint64_t)blockIdx.x*OBJECT_GROUP+lid, N = cells(L);
	float vx = 0.0f, vy = 0.0f, vz = 0.0f; uint32_t c = 0u;
	if(n<N && L.flags[n]==flag_marker) {
		const uint64_t plane = (uint64_t)L.Nx*L.Ny, r = n%plane;
		const Float3 p = cell_position(L, (uint32_t)(r%L.Nx), (uint32_t)(r/L.Nx), (uint32_t)(n/plane));
		if constexpr(KIND==0) { vx = p.x; vy = p.y; vz = p.z; c = 1u; }
		else if constexpr(KIND==1) { vx = L.F[n]; vy = L.F[N+n]; vz = L.F[2ull*N+n]; }
		else { const Float3 t = cross3(p-f3(cx, cy, cz), f3(L.F[n], L.F[N+n], L.F[2ull*N+n])); vx = t.x; vy = t.y; vz = t.z; }
	}
	cache[0][lid] = vx; cache[1][lid] = vy; cache[2][lid] = vz; count[lid] = c;
	__syncthreads();
	for(uint32_t s=1u; s<OBJECT_GROUP; s*=2u) {
		if(lid%(2u*s)==0u) { cache[0][lid] += cache[0][lid+s]; cache[1][lid] += cache[1][lid+s]; cache[2][lid] += cache[2][lid+s]; count[lid] += count[lid+s]; }
		__syncthreads();
	}
	if(lid==0u) partial[blockIdx.x] = ObjectPartial{ cache[0][0], cache[1][0], cache[2][0], count[0] };
}
#if defined(FX3D_TU_LBM)
__global__ void __launch_bounds__(32) k_object_total(const ObjectPartial* partial, const uint32_t groups, const uint32_t kind, float* object_sum) {
	const uint32_t lane = threadIdx.x;
	float sx = 0.0f, sy = 0.0f, sz = 0.0f; uint32_t cells_total = 0u;
	for(uint32_t base=0u; base<groups; base+=32u) {
		ObjectPartial p = ObjectPartial{ 0.0f, 0.0f, 0.0f, 0u };
		if(base+lane<groups) p = partial[base+lane];
		const bool any = kind==0u ? p.cells>0u : (p.x!=0.0f || p.y!=0.0f || p.z!=0.0f);
		uint32_t todo = __ballot_sync(0xFFFFFFFFu, any);
		while(todo) { // ascending group order; adding a zero component changes nothing, so components need no separate test
#if defined(FX3D_HOST_EMULATION)
			const int k = __builtin_ctz(todo);
#else
			const int k = __ffs((int)todo)-1;
#endif
			todo &= todo-1u;
			sx += __uint_as_float(__shfl_sync(0xFFFFFFFFu, __float_as_uint(p.x), k)); sy += __uint_as_float(__shfl_sync(0xFFFFFFFFu, __float_as_uint(p.y), k));
			sz += __uint_as_float(__shfl_sync(0xFFFFFFFFu, __float_as_uint(p.z), k)); cells_total += __shfl_sync(0xFFFFFFFFu, p.cells, k);
		}
	}
	if(lane==0u) { object_sum[0] = sx; object_sum[1] = sy; object_sum[2] = sz; object_sum[3] = __uint_as_float(cells_total); }
}
// transfer_extract_F / transfer__insert_F, :2173-2196: three float planes per side
template<bool EXTRACT>
__global__ void __launch_bounds__(128) k_transfer_F(const Lattice L, const uint32_t axis, float* buf_p, float* buf_m) {
	const uint32_t a = blockIdx.x*blockDim.x+threadIdx.x, A = face_area_of(L, axis), len = axis==0u ? L.Nx : axis==1u ? L.Ny : L.Nz;
	if(a>=A) return;
	const uint64_t N = cells(L);
	uint32_t x, y, z;
	for(int side=0; side<2; side++) {
		face_cell_of(L, axis, a, side==0 ? (EXTRACT ? len-2u : len-1u) : (EXTRACT ? 1u : 0u), x, y, z);
		const uint64_t n = lin(L, x, y, z);
		float* fb = side==0 ? buf_p : buf_m;
		for(uint32_t k=0u; k<3u; k++) { if(EXTRACT) fb[(uint64_t)k*A+a] = L.F[k*N+n]; else L.F[k*N+n] = fb[(uint64_t)k*A+a]; }
	}
}
// direct peer pull of the F halo (communicate_F, src/lbm.cpp:1395-1397) and of the flags halo alone (communicate_flags, :1391-1393;
// transfer_extract/insert_flags, src/kernel.cpp:2160-2171: one byte per face cell)
__global__ void __launch_bounds__(128) k_exchange_F(const Lattice L, const uint32_t axis, const float* F_plus, const float* F_minus) {
	const uint32_t a = blockIdx.x*blockDim.x+threadIdx.x, A = face_area_of(L, axis), len = axis==0u ? L.Nx : axis==1u ? L.Ny : L.Nz;
	if(a>=A) return;
	const uint64_t N = cells(L);
	uint32_t x, y, z;
	for(int side=0; side<2; side++) {
		const float* src = side==0 ? F_plus : F_minus;
		face_cell_of(L, axis, a, side==0 ? 1u : len-2u, x, y, z);
		const uint64_t ns = lin(L, x, y, z);
		face_cell_of(L, axis, a, side==0 ? len-1u : 0u, x, y, z);
		const uint64_t nd = lin(L, x, y, z);
		for(uint32_t k=0u; k<3u; k++) L.F[k*N+nd] = src[k*N+ns];
	}
}
__global__ void __launch_bounds__(128) k_exchange_flags(const Lattice L, const uint32_t axis, const uint8_t* flags_plus, const uint8_t* flags_minus) {
	const uint32_t a = blockIdx.x*blockDim.x+threadIdx.x, A = face_area_of(L, axis), len = axis==0u ? L.Nx : axis==1u ? L.Ny : L.Nz;
	if(a>=A) return;
	uint32_t x, y, z;
	for(int side=0; side<2; side++) {
		const uint8_t* src = side==0 ? flags_plus : flags_minus;
		face_cell_of(L, axis, a, side==0 ? 1u : len-2u, x, y, z);
		const uint64_t ns = lin(L, x, y, z);
		face_cell_of(L, axis, a, side==0 ? len-1u : 0u, x, y, z);
		L.flags[lin(L, x, y, z)] = src[ns];
	}
}
template<bool EXTRACT>
__global__ void __launch_bounds__(128) k_transfer_flags(const Lattice L, const uint32_t axis, uint8_t* buf_p, uint8_t* buf_m) {
	const uint32_t a = blockIdx.x*blockDim.x+threadIdx.x, A = face_area_of(L, axis), len = axis==0u ? L.Nx : axis==1u ? L.Ny : L.Nz;
	if(a>=A) return;
	uint32_t x, y, z;
	for(int side=0; side<2; side++) {
		face_cell_of(L, axis, a, side==0 ? (EXTRACT ? len-2u : len-1u) : (EXTRACT ? 1u : 0u), x, y, z);
		const uint64_t n = lin(L, x, y, z);
		uint8_t* b = side==0 ? buf_p : buf_m;
		if(EXTRACT) b[a] = L.flags[n]; else L.flags[n] = b[a];
	}
}
#endif // FX3D_TU_LBM

// ================================================================================================================
// halo transfer. Direction lists per face side (position b pairs opposite directions on the two sides),
// src/kernel.cpp:2069-2101; face-cell decomposition of a per axis :2053-2068.
// ================================================================================================================
template<int Q> FX3D_HDC constexpr int transfers() { return Q==19 ? 5 : 9; }
template<int Q> FX3D_HDC constexpr int xfer_dir(int side, int b) {
	if constexpr(Q==19) {
		constexpr uint8_t t[6][5] = { { 1, 7,13, 9,15 }, { 2, 8,14,10,16 }, { 3, 7,14,11,17 }, { 4, 8,13,12,18 }, { 5, 9,16,11,18 }, { 6,10,15,12,17 } };
		return t[side][b];
	} else {
		constexpr uint8_t t[6][9] = { { 1, 7,13, 9,15,19,26,21,23 }, { 2, 8,14,10,16,20,25,22,24 }, { 3, 7,14,11,17,19,24,21,25 },
			{ 4, 8,13,12,18,20,23,22,26 }, { 5, 9,16,11,18,19,22,23,25 }, { 6,10,15,12,17,20,21,24,26 } };
		return t[side][b];
	}
}
FX3D_HD uint32_t face_area(const Lattice& L, uint32_t axis) { return axis==0u ? L.Ny*L.Nz : axis==1u ? L.Nz*L.Nx : L.Nx*L.Ny; }
FX3D_HD uint32_t axis_len(const Lattice& L, uint32_t axis) { return axis==0u ? L.Nx : axis==1u ? L.Ny : L.Nz; }
FX3D_HD void face_coords(const Lattice& L, uint32_t axis, uint32_t a, uint32_t layer, uint32_t& x, uint32_t& y, uint32_t& z) {
	if(axis==0u) { x = layer; y = a%L.Ny; z = a/L.Ny; }
	else if(axis==1u) { x = a/L.Nz; y = layer; z = a%L.Nz; }
	else { x = a%L.Nx; y = a/L.Nx; z = layer; }
}
// physical offset (slot included) that extract_fi() reads for direction i at face cell (x,y,z), src/kernel.cpp:2102-2110
FX3D_HD uint64_t extract_addr(const Lattice& L, uint32_t odd, int i, uint32_t x, uint32_t y, uint32_t z) {
	const uint32_t slot = odd ? ((i&1) ? i+1 : i-1) : i;
	if(i&1) { x = step_rt(dir_x(i), x, L.Nx); y = step_rt(dir_y(i), y, L.Ny); z = step_rt(dir_z(i), z, L.Nz); } // cell j[i]
	return (uint64_t)slot*L.slot+phys(L, x, y, z);
}
// physical offset that insert_fi() writes for direction i at face cell (x,y,z), src/kernel.cpp:2111-2119
FX3D_HD uint64_t insert_addr(const Lattice& L, uint32_t odd, int i, uint32_t x, uint32_t y, uint32_t z) {
	const uint32_t slot = odd ? i : ((i&1) ? i+1 : i-1);
	if(!(i&1)) { x = step_rt(dir_x(i-1), x, L.Nx); y = step_rt(dir_y(i-1), y, L.Ny); z = step_rt(dir_z(i-1), z, L.Nz); } // cell j[i-1]
	return (uint64_t)slot*L.slot+phys(L, x, y, z);
}

// staged transfer through linear buffers laid out buf[b*A+a] like the reference's (kept for interface parity and as
// the building block of host-staged exchange): EXTRACT=true transfer_extract_fi, false transfer__insert_fi
template<int Q, int ST, bool EXTRACT>
__global__ void __launch_bounds__(128) k_transfer_fi(const Lattice L, const uint32_t axis, void* buf_p, void* buf_m) {
	// one thread per (face cell a, transferred direction b = blockIdx.y); the loads of both sides are issued before the stores
	typedef typename Codec<ST>::elem_t E;
	const uint32_t a = blockIdx.x*blockDim.x+threadIdx.x, A = face_area(L, axis), len = axis_len(L, axis);
	if(a>=A) return;
	const int b = (int)blockIdx.y;
	E* fi = reinterpret_cast<E*>(L.fi);
	E* bp = reinterpret_cast<E*>(buf_p); E* bm = reinterpret_cast<E*>(buf_m);
	uint32_t xp, yp, zp, xm, ym, zm;
	face_coords(L, axis, a, EXTRACT ? len-2u : len-1u, xp, yp, zp);
	face_coords(L, axis, a, EXTRACT ? 1u : 0u, xm, ym, zm);
	const int ip = xfer_dir<Q>(2*(int)axis, b), im = xfer_dir<Q>(2*(int)axis+1, b);
	const uint64_t k = (uint64_t)b*A+a;
	if(EXTRACT) {
		const E vp = fi[extract_addr(L, L.odd, ip, xp, yp, zp)], vm = fi[extract_addr(L, L.odd, im, xm, ym, zm)];
		bp[k] = vp; bm[k] = vm;
	} else {
		const E vp = bp[k], vm = bm[k];
		fi[insert_addr(L, L.odd, ip, xp, yp, zp)] = vp; fi[insert_addr(L, L.odd, im, xm, ym, zm)] = vm; // (writing the halo cells' whole 32-byte sectors instead changed nothing: 119 us either way for two 512^2 faces -- the kernel is bound by the 32 separate transactions of every warp access, not by partial sectors)
	}
}
// rho/u/flags halo: 4 float planes then one byte plane at byte 16*A, src/kernel.cpp:2133-2158
template<bool EXTRACT>
__global__ void __launch_bounds__(128) k_transfer_rho_u_flags(const Lattice L, const uint32_t axis, void* buf_p, void* buf_m) {
	const uint32_t a = blockIdx.x*blockDim.x+threadIdx.x, A = face_area(L, axis), len = axis_len(L, axis);
	if(a>=A) return;
	const uint64_t N = cells(L);
	uint32_t x, y, z;
	for(int side=0; side<2; side++) {
		face_coords(L, axis, a, side==0 ? (EXTRACT ? len-2u : len-1u) : (EXTRACT ? 1u : 0u), x, y, z);
		const uint64_t n = lin(L, x, y, z);
		float* fb = reinterpret_cast<float*>(side==0 ? buf_p : buf_m);
		uint8_t* cb = reinterpret_cast<uint8_t*>(side==0 ? buf_p : buf_m);
		if(EXTRACT) {
			fb[a] = L.rho[n]; fb[A+a] = L.u[n]; fb[2ull*A+a] = L.u[N+n]; fb[3ull*A+a] = L.u[2ull*N+n]; cb[16ull*A+a] = L.flags[n];
		} else {
			L.rho[n] = fb[a]; L.u[n] = fb[A+a]; L.u[N+n] = fb[2ull*A+a]; L.u[2ull*N+n] = fb[3ull*A+a]; L.flags[n] = cb[16ull*A+a];
		}
	}
}

// Direct peer exchange of one axis: every face cell pulls, over NVLink peer loads (or the same GPU), exactly the
// raw DDF bits that the reference's extract -> host swap -> insert sequence would have delivered
// (src/lbm.cpp:1355-1383: domain d's "+" buffer goes to its +axis neighbour's "-" side and vice versa).
//   my layer len-1 ("+" insert, list p)  <-  +axis neighbour's "-" extract at its layer 1   (list m)
//   my layer 0     ("-" insert, list m)  <-  -axis neighbour's "+" extract at its layer len-2 (list p)
// All domains have identical local geometry, so the neighbour's addresses follow from my Lattice.
template<int Q, int ST>
__global__ void __launch_bounds__(128) k_exchange_fi(const Lattice L, const uint32_t axis, const void* fi_plus, const void* fi_minus) {
	// one thread per (face cell a, transferred direction b = blockIdx.y): two independent peer loads in flight per thread
	typedef typename Codec<ST>::elem_t E;
	const uint32_t a = blockIdx.x*blockDim.x+threadIdx.x, A = face_area(L, axis), len = axis_len(L, axis);
	if(a>=A) return;
	const int b = (int)blockIdx.y;
	E* fi = reinterpret_cast<E*>(L.fi);
	const E* fp = reinterpret_cast<const E*>(fi_plus); const E* fm = reinterpret_cast<const E*>(fi_minus);
	const int ip = xfer_dir<Q>(2*(int)axis, b), im = xfer_dir<Q>(2*(int)axis+1, b);
	uint32_t xi, yi, zi, xe, ye, ze;
	face_coords(L, axis, a, 1u, xe, ye, ze);
	const E vp = fp[extract_addr(L, L.odd, im, xe, ye, ze)]; // what the +neighbour sends towards -axis lands in my +halo
	face_coords(L, axis, a, len-2u, xe, ye, ze);
	const E vm = fm[extract_addr(L, L.odd, ip, xe, ye, ze)];
	face_coords(L, axis, a, len-1u, xi, yi, zi);
	fi[insert_addr(L, L.odd, ip, xi, yi, zi)] = vp;
	face_coords(L, axis, a, 0u, xi, yi, zi);
	fi[insert_addr(L, L.odd, im, xi, yi, zi)] = vm;
}
// y and z faces: for a fixed face row and direction the transferred elements are one whole x-row of one slot, and the row the
// neighbour extracts from and the row I insert into carry the same x shift -- so the exchange of a (row, direction, side) is a
// copy of the full pitch row, done with aligned 16-byte vectors (512 bytes per warp request over NVLink).
// grid = (face rows, transfers, 2 sides), any 1-D block.
template<int Q, int ST>
__global__ void __launch_bounds__(128) k_exchange_fi_rows(const Lattice L, const uint32_t axis, const void* fi_plus, const void* fi_minus) {
	typedef typename Codec<ST>::elem_t E;
	const uint32_t r = blockIdx.x, len = axis_len(L, axis);
	const int b = (int)blockIdx.y;
	const bool plus = blockIdx.z==0u;
	const int ip = xfer_dir<Q>(2*(int)axis, b), im = xfer_dir<Q>(2*(int)axis+1, b);
	// a face cell of this row (x = 0): the rows of its source and destination elements are the rows of everything in the row
	uint32_t xe = 0u, ye = axis==1u ? (plus ? 1u : len-2u) : r, ze = axis==1u ? r : (plus ? 1u : len-2u);
	uint32_t xi = 0u, yi = axis==1u ? (plus ? len-1u : 0u) : r, zi = axis==1u ? r : (plus ? len-1u : 0u);
	const uint64_t src = extract_addr(L, L.odd, plus ? im : ip, xe, ye, ze), dst = insert_addr(L, L.odd, plus ? ip : im, xi, yi, zi);
	// the cell row that holds the element: its pitch row, and with an x halo (where the cell x=0 is the LAST element of a pitch row and x=1 the first of
	// the next one) the Nx elements from that last element on; the addresses above may have been stepped from x=0 to x=1 or x=Nx-1, which lie one pitch row later
	const uint64_t src_pitch = (src-src%L.px)-((L.Hx && src%L.px!=L.px-1u) ? L.px : 0u), dst_pitch = (dst-dst%L.px)-((L.Hx && dst%L.px!=L.px-1u) ? L.px : 0u);
	const E* sp = reinterpret_cast<const E*>(plus ? fi_plus : fi_minus)+src_pitch+L.xo;
	E* dp = reinterpret_cast<E*>(L.fi)+dst_pitch+L.xo;
	const uint32_t n = L.Hx ? L.Nx : L.px; // elements to copy (source and destination have the same alignment: the domains share one layout)
	uint32_t head = (uint32_t)(((16u-(uint32_t)(reinterpret_cast<uintptr_t>(sp)&15u))&15u)/sizeof(E));
	if(head>n) head = n;
	const uint32_t nvec = (n-head)*(uint32_t)sizeof(E)/16u, tail = head+nvec*(16u/(uint32_t)sizeof(E));
	for(uint32_t k=threadIdx.x; k<head; k+=blockDim.x) dp[k] = sp[k];
	const uint4* sv = reinterpret_cast<const uint4*>(sp+head);
	uint4* dv = reinterpret_cast<uint4*>(dp+head);
	for(uint32_t v=threadIdx.x; v<nvec; v+=blockDim.x) dv[v] = sv[v];
	for(uint32_t k=tail+threadIdx.x; k<n; k+=blockDim.x) dp[k] = sp[k];
}
struct PeerFields { const float* rho; const float* u; const uint8_t* flags; };
#if defined(FX3D_TU_LBM) // non-template kernels are defined in exactly one translation unit
__global__ void __launch_bounds__(128) k_exchange_rho_u_flags(const Lattice L, const uint32_t axis, const PeerFields plus, const PeerFields minus) {
	const uint32_t a = blockIdx.x*blockDim.x+threadIdx.x, A = face_area(L, axis), len = axis_len(L, axis);
	if(a>=A) return;
	const uint64_t N = cells(L);
	uint32_t x, y, z;
	for(int side=0; side<2; side++) {
		const PeerFields& src = side==0 ? plus : minus;
		face_coords(L, axis, a, side==0 ? 1u : len-2u, x, y, z);
		const uint64_t ns = lin(L, x, y, z);
		face_coords(L, axis, a, side==0 ? len-1u : 0u, x, y, z);
		const uint64_t nd = lin(L, x, y, z);
		L.rho[nd] = src.rho[ns]; L.u[nd] = src.u[ns]; L.u[N+nd] = src.u[N+ns]; L.u[2ull*N+nd] = src.u[2ull*N+ns]; L.flags[nd] = src.flags[ns];
	}
}

#endif // FX3D_TU_LBM

// storage codec self-test: out[k] = encode(in[k]) with the fast and with the literal formula, dec[h] likewise
template<int ST>
__global__ void k_codec_encode(const float* in, uint16_t* out, uint64_t n) {
	const uint64_t k = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x;
	if(k<n) out[k] = Codec<ST>::encode(in[k]);
}
template<int ST>
__global__ void k_codec_decode(const uint16_t* in, float* out, uint64_t n) {
	const uint64_t k = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x;
	if(k<n) out[k] = Codec<ST>::decode(in[k]);
}
#if defined(FX3D_TU_RUNTIME)
// packed-lane self-test: moments, equilibrium and the full cell update in F2 (two cells per operation) against the scalar
// instantiation of the same templates, bit for bit, on pseudo-random populations (guards against instruction contraction)
template<int Q, int COLL, bool VF> __global__ void k_selftest_lanes(unsigned long long samples_per_thread, float S, unsigned long long* mismatches) {
	unsigned long long s = ((unsigned long long)blockIdx.x*blockDim.x+threadIdx.x)*0x9E3779B97F4A7C15ull+0x7654321ull, bad = 0ull;
	auto next = [&]() { s ^= s<<13; s ^= s>>7; s ^= s<<17; return (uint32_t)(s>>16); };
	auto rnd = [&]() { return (float)(next()&0xFFFFFFu)/16777216.0f-0.5f; };
	const float inv = 1.0f/S;
	for(unsigned long long k=0ull; k<samples_per_thread; k++) {
		float fa[Q], fb[Q]; F2 f2[Q];
		const float amp = (k&3ull)==0ull ? 0.0f : 0.02f; // every fourth sample is a fluid at rest (all-zero populations)
		static_for<0, Q, 1>([&](auto I) { fa[I] = amp*rnd()*S; fb[I] = amp*rnd()*S; if(S!=1.0f) { fa[I] = rintf(fa[I]); fb[I] = rintf(fb[I]); } f2[I] = make_f2(fa[I], fb[I]); });
		float ra, uxa, uya, uza, rb, uxb, uyb, uzb; F2 r2, ux2, uy2, uz2;
		moments<Q, float>(fa, S, inv, ra, uxa, uya, uza); moments<Q, float>(fb, S, inv, rb, uxb, uyb, uzb); moments<Q, F2>(f2, S, inv, r2, ux2, uy2, uz2);
		auto ne = [](float x, float y) { return __float_as_uint(x)!=__float_as_uint(y) ? 1ull : 0ull; };
		bad += ne(f2_lo(r2), ra)+ne(f2_hi(r2), rb)+ne(f2_lo(ux2), uxa)+ne(f2_hi(ux2), uxb)+ne(f2_lo(uy2), uya)+ne(f2_hi(uy2), uyb)+ne(f2_lo(uz2), uza)+ne(f2_hi(uz2), uzb);
		float ea[Q], eb[Q]; F2 e2[Q];
		equilibrium<Q, float>(ra, uxa, uya, uza, S, ea); equilibrium<Q, float>(rb, uxb, uyb, uzb, S, eb); equilibrium<Q, F2>(r2, ux2, uy2, uz2, S, e2);
		static_for<0, Q, 1>([&](auto I) { bad += ne(f2_lo(e2[I]), ea[I])+ne(f2_hi(e2[I]), eb[I]); });
		float o0, o1, o2, o3; F2 p0, p1, p2, p3;
		const bool e_lo = (k&7ull)==5ull, e_hi = (k&15ull)==9ull;
		collide_cell<Q, COLL, VF, float>(fa, S, inv, e_lo, false, 1.01f, 0.02f, -0.03f, 0.04f, 1e-4f, -2e-4f, 3e-4f, 1.7f, o0, o1, o2, o3);
		collide_cell<Q, COLL, VF, float>(fb, S, inv, e_hi, false, 0.99f, -0.01f, 0.05f, 0.0f, 1e-4f, -2e-4f, 3e-4f, 1.7f, o0, o1, o2, o3);
		collide_cell<Q, COLL, VF, F2>(f2, S, inv, e_lo, e_hi, make_f2(1.01f, 0.99f), make_f2(0.02f, -0.01f), make_f2(-0.03f, 0.05f), make_f2(0.04f, 0.0f), 1e-4f, -2e-4f, 3e-4f, 1.7f, p0, p1, p2, p3);
		static_for<0, Q, 1>([&](auto I) { bad += ne(f2_lo(f2[I]), fa[I])+ne(f2_hi(f2[I]), fb[I]); });
	}
	if(bad) atomicAdd(mismatches, bad);
}
// division self-test: the shared-reciprocal division of lbm_core.cuh against operator/ on pseudo-random and edge operands
__global__ void k_selftest_division(unsigned long long samples_per_thread, unsigned long long* mismatches) {
	unsigned long long s = ((unsigned long long)blockIdx.x*blockDim.x+threadIdx.x)*0x9E3779B97F4A7C15ull+0x1234567ull, bad = 0ull;
	auto next = [&]() { s ^= s<<13; s ^= s>>7; s ^= s<<17; return (uint32_t)(s>>16); };
	for(unsigned long long k=0ull; k<samples_per_thread; k++) {
		const uint32_t r0 = next(), r1 = next(), r2 = next(), r3 = next(), mode = next()&7u;
		// denominators: rho-like (0.5..2), scaled rho (2^14..2^16), or any exponent; numerators: tiny..moderate, zeros, denormals
		const uint32_t eb = mode<4u ? 126u+(r0&1u) : mode<6u ? 141u+(r0&1u) : (r0>>8)&0xFFu;
		const float b = __uint_as_float((r1&0x807FFFFFu)|(eb<<23));
		const uint32_t ea = (mode&1u) ? 100u+(r2%60u) : (r2>>3)&0xFFu;
		float a0 = __uint_as_float((r3&0x807FFFFFu)|(ea<<23)), a1 = __uint_as_float((next()&0x807FFFFFu)|(((ea+r0)%255u)<<23)), a2 = (mode==3u) ? 0.0f : __uint_as_float(next());
		if(a0!=a0||a1!=a1||a2!=a2||b!=b) continue;
		float q0, q1, q2;
		vdiv3(a0, a1, a2, b, q0, q1, q2);
		const float e0 = a0/b, e1 = a1/b, e2 = a2/b;
		bad += (__float_as_uint(q0)!=__float_as_uint(e0) && e0==e0)+(__float_as_uint(q1)!=__float_as_uint(e1) && e1==e1)+(__float_as_uint(q2)!=__float_as_uint(e2) && e2==e2);
		F2 p0, p1, p2;
		vdiv3(make_f2(a0, a1), make_f2(a1, a2), make_f2(a2, a0), make_f2(b, b), p0, p1, p2);
		bad += (__float_as_uint(f2_lo(p0))!=__float_as_uint(e0) && e0==e0)+(__float_as_uint(f2_hi(p0))!=__float_as_uint(e1) && e1==e1)+(__float_as_uint(f2_hi(p1))!=__float_as_uint(e2) && e2==e2);
		const float h = vdiv1(0.5f, b), eh = 0.5f/b;
		bad += (__float_as_uint(h)!=__float_as_uint(eh) && eh==eh);
	}
	if(bad) atomicAdd(mismatches, bad);
}
// exhaustive check over all 2^32 binary32 inputs: counts inputs where the one-multiply FP16C encode differs from
// the reference's literal formula (and all 2^16 codes for decode); NaN payloads excluded for decode comparisons
__global__ void k_fp16c_exhaustive(unsigned long long* mismatches, uint32_t* first_bad) {
	const uint64_t stride = (uint64_t)gridDim.x*blockDim.x;
	unsigned long long bad = 0ull;
	for(uint64_t k = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x; k<(1ull<<32); k += stride) {
		const float x = __uint_as_float((uint32_t)k);
		if(x!=x) continue; // NaN inputs are outside the contract
		if(fp16c_encode(x)!=fp16c_encode_literal(x)) { if(bad==0ull) first_bad[0] = (uint32_t)k; bad++; }
		if(k<65536ull) {
			const float a = fp16c_decode((uint16_t)k), b = fp16c_decode_literal((uint16_t)k);
			if(__float_as_uint(a)!=__float_as_uint(b)) { if(bad==0ull) first_bad[0] = (uint32_t)k; bad++; }
		}
	}
	if(bad) {
#if defined(FX3D_HOST_EMULATION)
		*mismatches += bad;
#else
		atomicAdd(mismatches, bad);
#endif
	}
}
#endif // FX3D_TU_RUNTIME

} // namespace fx3d
