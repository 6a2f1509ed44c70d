// fx3d_internal.cuh -- shared host-side plumbing of libfx3d_cuda: error reporting, launch macro, layout rules.
#pragma once
#include "../../include/fx3d.h"
#include "lbm_kernels.cuh"
#include <atomic>
#include <cstdio>
#include <string>
#include <vector>

namespace fx3d {

void set_error(const std::string& msg);       // stores the thread-local text returned by fx3d_last_error()
extern std::atomic<uint64_t> g_launches;      // kernels launched so far (fx3d_launch_count)
extern std::atomic<int> g_variant;            // 0 auto, 1 general one-cell-per-thread kernel, 2 / 4 vector kernel with that many cells per thread

#if defined(FX3D_HOST_EMULATION)
// test-only build (tests/emul): kernels run as OS threads on host pointers; there is no device to select
#define FX3D_LAUNCH(kernel, grid, block, stream, ...) do { ::fx3d::g_launches++; ::emul::launch((grid), (block), [&]() { kernel(__VA_ARGS__); }); } while(0)
#define FX3D_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) do { ::fx3d::g_launches++; ::emul::launch((grid), (block), [&]() { kernel(__VA_ARGS__); }, (smem)); } while(0)
inline int use_device(int) { return FX3D_OK; }
inline int check_launch(const char*) { return FX3D_OK; }
#else
#define FX3D_LAUNCH(kernel, grid, block, stream, ...) do { ::fx3d::g_launches++; kernel<<<(grid), (block), 0, (cudaStream_t)(stream)>>>(__VA_ARGS__); } while(0)
#define FX3D_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) do { ::fx3d::g_launches++; kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__); } while(0)
int use_device(int device);                   // cudaSetDevice with error translation
int check_launch(const char* what);           // cudaGetLastError with error translation
int cuda_fail(cudaError_t e, const char* what);
#endif

// DDF layout rules (see lbm_kernels.cuh). Without an x halo: rows of Nx elements, pitch rounded up to 64 elements. With an x halo the NON-HALO cells
// of a row (x = 1 .. Nx-2) start on a pitch boundary: x = 1 is element 0 of pitch row r+1, the halo cell x = Nx-1 follows the row, and the halo cell
// x = 0 is the last element of pitch row r (xo = px-1; the slot carries one extra pitch row). For rows of 512 cells and more the pitch is a multiple of
// 512 elements, so that every row of non-halo cells starts on a multiple of its own length in bytes -- measured on B200 (profiles/r02_row_kernel_tuning.txt):
// bulk copies of 1-KB rows that start 128 bytes into a 1280-byte pitch (the first layout: xo = 63) ran the collide kernel of a 514x512x512 domain at
// 2.08 ms (FP16S) / 3.55 ms (FP32); the same rows on 2-KB boundaries 1.89 / 3.37 ms (periodic 512^3: 1.73 / 3.32 ms). The price is memory: an
// x-decomposed 514-cell row occupies 1024 elements.
inline bool make_lattice(const fx3d_lattice* in, uint64_t t, float fx, float fy, float fz, Lattice& L) {
	if(!in) { set_error("lattice is null"); return false; }
	if(in->Nx==0u||in->Ny==0u||in->Nz==0u) { set_error("lattice size is 0"); return false; }
	if(in->velocity_set!=19u&&in->velocity_set!=27u) { set_error("velocity_set must be 19 or 27 (D3Q19 / D3Q27)"); return false; }
	if(in->collision>1u||in->storage>2u||in->Dx==0u||in->Dy==0u||in->Dz==0u) { set_error("invalid collision/storage/domain count"); return false; }
	L.Nx = in->Nx; L.Ny = in->Ny; L.Nz = in->Nz;
	L.Hx = in->Dx>1u; L.Hy = in->Dy>1u; L.Hz = in->Dz>1u;
	if((L.Hx&&in->Nx<3u)||(L.Hy&&in->Ny<3u)||(L.Hz&&in->Nz<3u)) { set_error("a decomposed axis needs at least 3 cells (halo + 1 + halo)"); return false; }
	const uint32_t align = (L.Hx && in->Nx-2u>=512u) ? 512u : 64u;
	L.px = ((in->Nx+align-1u)/align)*align;
	L.xo = L.Hx ? L.px-1u : 0u;
	L.slot = (uint64_t)L.px*((uint64_t)in->Ny*in->Nz+(L.Hx ? 1ull : 0ull));
	if(L.slot>0xFFFFFFFFull) { set_error("a domain may hold at most 2^32-1 (padded) cells"); return false; }
	L.slot32 = (uint32_t)L.slot;
	L.fi = in->fi; L.rho = in->rho; L.u = in->u; L.flags = in->flags;
	L.w = in->w; L.fx = fx; L.fy = fy; L.fz = fz;
	L.odd = (uint32_t)(t&1ull);
	L.eb = (in->features&FX3D_EQUILIBRIUM_BOUNDARIES) ? 1u : 0u;
	L.upd = (in->features&FX3D_UPDATE_FIELDS) ? 1u : 0u;
	L.mb = (in->features&FX3D_MOVING_BOUNDARIES) ? 1u : 0u;
	L.F = (in->features&FX3D_FORCE_FIELD) ? in->F : nullptr; // entry points that dereference it check for null themselves (fx3d_fi_bytes etc. run before the buffers exist)
	return true;
}
inline size_t elem_bytes(uint32_t storage) { return storage==FX3D_FP32 ? 4u : 2u; }

// stream_collide instantiations live in one translation unit per (velocity set, storage), see sc_inst.cu
extern std::atomic<uint64_t> g_kind_launches[8]; // stream_collide launches by kernel kind: 0 general, 1 vector, 2 cp.async ring, 3 bulk copies (rows), 4 bulk copies (segments), 5 bulk loads + direct stores, 6 one cell per thread at high occupancy, 7 bulk copies (rows, shared ring)
inline uint32_t pipe_cells_of(uint32_t velocity_set, uint32_t storage) { return (storage==FX3D_FP32 && velocity_set>19u) ? 2u : 4u; } // = pipe_cells<Q,ST>()
template<int Q, int ST> int launch_stream_collide(const Lattice& L, const Region& R, int cells_per_thread, int collision, bool volume_force, void* stream, int reserve=0, int ext=0, const RowPeers* fused=nullptr); // cells_per_thread 0: persistent kernel chosen automatically (-1: cp.async ring, -2: whole-row bulk copies or nothing, -3: bulk copies wherever eligible), leaving `reserve` resident-block slots free

} // namespace fx3d
