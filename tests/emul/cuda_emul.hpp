// cuda_emul.hpp -- a minimal host-side emulation of the CUDA execution model, TEST INFRASTRUCTURE ONLY.
//
// The development container has no GPU. So that the *logic* of the product kernels (addressing, warp shuffles,
// pack/unpack, operation order) can be checked against the CPU oracle before spending GPU time, the kernel
// sources under fluidx3d_b200/csrc are also compilable as plain C++ with -DFX3D_HOST_EMULATION, in which case this
// header supplies threadIdx/blockIdx, warp collectives, the few intrinsics used, and a launcher that runs every
// CUDA thread of a block as an OS thread (warp collectives synchronise through std::barrier).
// The result (tests/_build/libfx3d_emul.so) is loaded by tests only; the product library is built by nvcc only
// and contains no CPU path.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cfenv>
#include <thread>
#include <vector>
#include <barrier>
#include <memory>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)

struct dim3 { unsigned x, y, z; dim3(unsigned x_=1, unsigned y_=1, unsigned z_=1) : x(x_), y(y_), z(z_) {} };
struct uint3_ { unsigned x, y, z; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{ x, y }; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{ x, y }; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{ x, y, z, w }; }
typedef void* cudaStream_t;

namespace emul {
struct WarpCtx {
	std::unique_ptr<std::barrier<>> bar;
	uint32_t slot[32];
	int pred[32];
	unsigned lanes;
};
struct ThreadCtx { WarpCtx* warp; unsigned lane; unsigned char* smem; std::barrier<>* block_bar; std::barrier<>* group_bar; };
inline thread_local ThreadCtx tctx;
}
inline thread_local uint3_ threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

namespace emul {
inline uint32_t exchange(uint32_t v, int src_lane, bool valid_src) {
	WarpCtx* w = tctx.warp;
	w->slot[tctx.lane] = v;
	w->bar->arrive_and_wait();
	const uint32_t r = (valid_src && src_lane>=0 && (unsigned)src_lane<w->lanes) ? w->slot[src_lane] : v;
	w->bar->arrive_and_wait();
	return r;
}
inline unsigned ballot(int p) {
	WarpCtx* w = tctx.warp;
	w->pred[tctx.lane] = p;
	w->bar->arrive_and_wait();
	unsigned m = 0u;
	for(unsigned l=0u; l<w->lanes; l++) if(w->pred[l]) m |= 1u<<l;
	w->bar->arrive_and_wait();
	return m;
}
inline unsigned char* block_smem() { return tctx.smem; }
inline void group_barrier(unsigned) { tctx.group_bar->arrive_and_wait(); } // the calling thread's 128-thread group (groups are consecutive thread ranges)
template<class F> void launch(dim3 grid, dim3 block, F&& body, size_t smem_bytes=0) {
	const unsigned nthreads = block.x*block.y*block.z, nwarps = (nthreads+31u)/32u;
	for(unsigned bz=0u; bz<grid.z; bz++) for(unsigned by=0u; by<grid.y; by++) for(unsigned bx=0u; bx<grid.x; bx++) {
		std::vector<WarpCtx> warps(nwarps);
		for(unsigned wi=0u; wi<nwarps; wi++) {
			warps[wi].lanes = std::min(32u, nthreads-32u*wi);
			warps[wi].bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)warps[wi].lanes);
		}
		std::vector<unsigned char> smem(smem_bytes+128);
		std::barrier<> block_bar((std::ptrdiff_t)nthreads);
		std::vector<std::unique_ptr<std::barrier<>>> group_bars; // named barriers of 128-thread groups (bar.sync id, 128)
		for(unsigned gi=0u; gi<(nthreads+127u)/128u; gi++) group_bars.push_back(std::make_unique<std::barrier<>>((std::ptrdiff_t)std::min(128u, nthreads-128u*gi)));
		std::vector<std::thread> th;
		th.reserve(nthreads);
		for(unsigned tid=0u; tid<nthreads; tid++) th.emplace_back([&, tid]() {
			threadIdx = uint3_{ tid%block.x, (tid/block.x)%block.y, tid/(block.x*block.y) };
			blockIdx = uint3_{ bx, by, bz };
			blockDim = block; gridDim = grid;
			tctx.warp = &warps[tid/32u]; tctx.lane = tid%32u; tctx.smem = smem.data()+((128-reinterpret_cast<uintptr_t>(smem.data())%128)%128);
			tctx.block_bar = &block_bar; tctx.group_bar = group_bars[tid/128u].get();
			body();
			tctx.group_bar->arrive_and_drop();
			tctx.warp->bar->arrive_and_drop(); // a thread that has returned no longer takes part in warp collectives
			block_bar.arrive_and_drop();       // ... nor in block barriers
		});
		for(auto& t : th) t.join();
	}
}
}

static inline void __syncthreads() { emul::tctx.block_bar->arrive_and_wait(); }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t __float_as_uint(float x) { uint32_t u; std::memcpy(&u, &x, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float x; std::memcpy(&x, &u, 4); return x; }
static inline uint32_t __shfl_down_sync(unsigned, uint32_t v, unsigned d) { return emul::exchange(v, (int)emul::tctx.lane+(int)d, emul::tctx.lane+d<32u); }
static inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned d) { return emul::exchange(v, (int)emul::tctx.lane-(int)d, emul::tctx.lane>=d); }
static inline float __shfl_down_sync(unsigned m, float v, unsigned d) { return __uint_as_float(__shfl_down_sync(m, __float_as_uint(v), d)); }
static inline float __shfl_up_sync(unsigned m, float v, unsigned d) { return __uint_as_float(__shfl_up_sync(m, __float_as_uint(v), d)); }
static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) { return emul::exchange(v, src, true); }
static inline unsigned __ballot_sync(unsigned, int p) { return emul::ballot(p); }
static inline int __any_sync(unsigned, int p) { return emul::ballot(p)!=0u; }

struct __half { uint16_t bits; };
static inline __half __float2half_rn(float x) { const _Float16 h = (_Float16)x; __half r; std::memcpy(&r.bits, &h, 2); return r; }
static inline float __half2float(__half h) { _Float16 v; std::memcpy(&v, &h.bits, 2); return (float)v; }
static inline uint16_t __half_as_ushort(__half h) { return h.bits; }
static inline __half __ushort_as_half(uint16_t u) { __half h; h.bits = u; return h; }
static inline float __fmul_rz(float a, float b) {
	const int old = std::fegetround();
	std::fesetround(FE_TOWARDZERO);
	volatile float va = a, vb = b;
	volatile float r = va*vb;
	std::fesetround(old);
	return r;
}
static inline float __fdiv_rn(float a, float b) { return a/b; }
