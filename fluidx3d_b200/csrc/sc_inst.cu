// sc_inst.cu -- stream_collide instantiations for one (velocity set, storage) pair; compiled once per pair with
// -DFX3D_Q=19|27 -DFX3D_ST=0|1|2 so that the six heavy translation units build in parallel.
#include <atomic>
#include <mutex>
#include <vector>
#include <utility>
#include "fx3d_internal.cuh"
#include <algorithm>

namespace fx3d {

static inline dim3 block_shape(uint32_t nx) { // 128 threads; x extent = smallest power of two covering the row, capped at 128
	uint32_t bx = 1u;
	while(bx<nx && bx<128u) bx <<= 1;
	return dim3(bx, 128u/bx, 1u);
}

// persistent pipelined kernel: grid = resident blocks only (SM count x blocks per SM by shared memory), each block walks its tiles
template<int Q, int COLL, int ST, bool VF, int ODD> static int launch_pipe_parity(const Lattice& L, const Region& R, const dim3& block, void* stream, int reserve) {
	const uint32_t tiles_x = (R.g1-R.g0+block.x-1u)/block.x, tiles_y = (R.y1-R.y0+block.y-1u)/block.y, nz = R.z1-R.z0;
	constexpr uint32_t smem = pipe_smem_bytes<Q, ST>();
	int sms = 148, per_sm = (int)std::max(1u, std::min((uint32_t)pipe_blocks_per_sm<Q, ST>(), (227u*1024u)/(smem+1024u)));
#if !defined(FX3D_HOST_EMULATION)
	static std::atomic<uint64_t> configured{0ull}; // per instantiation: bit d = the opt-in shared memory size is set on device d
	int dev = 0; cudaGetDevice(&dev);
	if(dev>=64 || !((configured.load()>>dev)&1ull)) {
		const cudaError_t e = cudaFuncSetAttribute(k_stream_collide_pipe<Q, COLL, ST, VF, ODD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if(e!=cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(stream_collide_pipe)");
		if(dev<64) configured.fetch_or(1ull<<dev);
	}
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#else
	sms = 2; per_sm = 1;
#endif
	const uint64_t all_blocks = (uint64_t)sms*(uint64_t)per_sm, blocks = all_blocks>2ull*(uint64_t)std::max(reserve, 0) ? all_blocks-(uint64_t)std::max(reserve, 0) : all_blocks, ntiles = (uint64_t)tiles_x*tiles_y*nz;
	if((uint64_t)tiles_x*tiles_y>0xFFFFFFFFull) { set_error("region has too many tile columns"); return FX3D_ERR_INVALID; }
	if(ntiles==0ull) return FX3D_OK;
	const dim3 grid((uint32_t)std::min<uint64_t>(ntiles, blocks), 1u, 1u); // every block takes an equal contiguous share of the tiles
	g_kind_launches[2]++;
	FX3D_LAUNCH_SMEM((k_stream_collide_pipe<Q, COLL, ST, VF, ODD>), grid, block, smem, stream, L, R, tiles_x, tiles_y);
	return check_launch("stream_collide (pipelined)");
}

template<int Q, int COLL, int ST, bool VF> static int launch_pipe(const Lattice& L, const Region& R, const dim3& block, void* stream, int reserve) {
	return L.odd ? launch_pipe_parity<Q, COLL, ST, VF, 1>(L, R, block, stream, reserve) : launch_pipe_parity<Q, COLL, ST, VF, 0>(L, R, block, stream, reserve);
}

// bulk-copy (TMA) kernel: the tile must span whole rows -- see the kernel's header comment. Block = (bx, T/bx) with bx = row cells / K, B blocks per
// SM; the ring depth S is whatever fits this block's share of the SM's shared memory (228 KB, 1 KB of it reserved per resident block), at least 2.
template<int Q, int ST> static uint32_t row_stages() {
	constexpr uint32_t stage = row_stage_bytes<Q, ST>(), B = (uint32_t)row_blocks<Q, ST>();
	constexpr uint32_t share = (228u*1024u)/B-1024u<227u*1024u ? (228u*1024u)/B-1024u : 227u*1024u;
	return std::min<uint32_t>(ROW_MAX_STAGES, (share-row_header<Q, ST>())/stage);
}
template<int Q, int ST> static dim3 row_block(const Lattice& L) { // (0,0,0) where the row does not divide into the block
	constexpr uint32_t K = (uint32_t)row_cells<Q, ST>(), T = row_threads<Q, ST>();
	const uint32_t inner = L.Nx-2u*L.Hx, bx = inner/K;
	if(inner%K!=0u || bx==0u || bx>T || T%bx!=0u) return dim3(0u, 0u, 0u);
	return dim3(bx, T/bx, 1u);
}
template<int Q, int ST> static bool tma_eligible(const Lattice& L, const Region& R) {
	const uint32_t esz = ST==ST_FP32 ? 4u : 2u, inner = L.Nx-2u*L.Hx;
	const dim3 block = row_block<Q, ST>(L);
	if(block.x==0u || R.g0!=0u || R.g1*4u!=inner || (inner*esz)%16u!=0u || (R.y1-R.y0)%block.y!=0u) return false;
	if(L.Hx ? (block.y!=1u || ((L.xo+1u)*esz)%16u!=0u) : L.xo!=0u) return false; // x halos: one-row tiles, the first non-halo cell starts a 16-byte chunk
	return row_stages<Q, ST>()>=2u;
}
// FX3D_ROW_DYNAMIC: the counter the blocks of one launch claim their tiles from -- one word per (device, stream), zeroed on the stream before every launch
// (launches on one stream are ordered, so one word per stream is enough)
static uint32_t* tile_counter_for(int dev, void* stream) {
#if defined(FX3D_HOST_EMULATION)
	static uint32_t word; (void)dev; (void)stream; word = 0u; return &word;
#else
	static std::mutex mu;
	static std::vector<std::pair<std::pair<int, void*>, uint32_t*>> table;
	std::lock_guard<std::mutex> lock(mu);
	for(auto& e : table) if(e.first.first==dev && e.first.second==stream) return e.second;
	uint32_t* p = nullptr;
	if(cudaMalloc(&p, 256u)!=cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
	table.push_back({ { dev, stream }, p });
	return p;
#endif
}
template<int Q, int COLL, int ST, bool VF, int ODD, bool SG = false, bool MB = false> static int launch_tma_parity(const Lattice& L, const Region& R, void* stream, int reserve, const RowPeers& peers) {
	const dim3 block = row_block<Q, ST>(L);
	const uint32_t tiles_y = (R.y1-R.y0)/block.y, nz = R.z1-R.z0;
	const uint32_t S = row_stages<Q, ST>();
	const uint32_t smem = row_header<Q, ST>()+S*row_stage_bytes<Q, ST>();
	int sms = 148, per_sm = row_blocks<Q, ST>();
#if !defined(FX3D_HOST_EMULATION)
	static std::atomic<uint64_t> configured{0ull}; // per instantiation: bit d = the opt-in shared memory size is set on device d
	int dev = 0; cudaGetDevice(&dev);
	if(dev>=64 || !((configured.load()>>dev)&1ull)) {
		const cudaError_t e = cudaFuncSetAttribute(k_stream_collide_tma<Q, COLL, ST, VF, ODD, SG, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if(e!=cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(stream_collide_tma)");
		if(dev<64) configured.fetch_or(1ull<<dev);
	}
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#else
	sms = 2; per_sm = 1;
#endif
	const uint64_t all_blocks = (uint64_t)sms*(uint64_t)per_sm, blocks = all_blocks>2ull*(uint64_t)std::max(reserve, 0) ? all_blocks-(uint64_t)std::max(reserve, 0) : all_blocks, ntiles = (uint64_t)tiles_y*nz;
	if(ntiles==0ull) return FX3D_OK;
	const dim3 grid((uint32_t)std::min<uint64_t>(ntiles, blocks), 1u, 1u);
	uint32_t* counter = nullptr;
#if FX3D_ROW_DYNAMIC
	if(ntiles>0xFFFF0000ull) { set_error("region has too many tiles"); return FX3D_ERR_INVALID; }
#if defined(FX3D_HOST_EMULATION)
	counter = tile_counter_for(0, stream);
#else
	counter = tile_counter_for(dev, stream);
#endif
	if(!counter) { set_error("stream_collide: no memory for the tile counter"); return FX3D_ERR_OUT_OF_MEMORY; }
#if !defined(FX3D_HOST_EMULATION)
	{ const cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), (cudaStream_t)stream); if(e!=cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(tile counter)"); }
#endif
#endif
	g_kind_launches[3]++;
	FX3D_LAUNCH_SMEM((k_stream_collide_tma<Q, COLL, ST, VF, ODD, SG, MB>), grid, block, smem, stream, L, R, tiles_y, S, peers, row_offsets<Q>(L, ST==ST_FP32 ? 4u : 2u), counter);
	return check_launch("stream_collide (bulk copies)");
}
// row segments (rows longer than 512 cells, shapes the whole-row kernel does not take): bulk loads, direct stores
template<int Q, int ST> static bool tmaseg_eligible(const Lattice& L, const Region& R, const dim3& block) {
	const uint32_t esz = ST==ST_FP32 ? 4u : 2u, groups = R.g1-R.g0;
	if((uint64_t)tmaseg_smem_bytes<Q, ST>()*tma_blocks_per_sm<Q, ST>()+2048u>227u*1024u) return false; // D3Q27 FP32
	if(block.x*block.y!=128u || block.y>4u || groups%block.x!=0u || (R.y1-R.y0)%block.y!=0u || block.x*4u*esz<32u) return false;
	if(((uint64_t)(L.Hx+R.g0*4u+L.xo)*esz)%16u!=0u) return false; // segment starts on a 16-byte boundary of its row
	if(L.Hx==0u && (L.Nx*esz)%16u!=0u) return false;               // so does the periodic wrap chunk
	return true;
}
template<int Q, int COLL, int ST, bool VF, int ODD, bool SG = false, bool MB = false> static int launch_hyb_parity(const Lattice& L, const Region& R, const dim3& block, void* stream, int reserve) {
	const uint32_t tiles_x = (R.g1-R.g0)/block.x, tiles_y = (R.y1-R.y0)/block.y, nz = R.z1-R.z0;
	constexpr uint32_t smem = tmaseg_smem_bytes<Q, ST>();
	int sms = 148, per_sm = tma_blocks_per_sm<Q, ST>();
#if !defined(FX3D_HOST_EMULATION)
	static std::atomic<uint64_t> configured{0ull};
	int dev = 0; cudaGetDevice(&dev);
	if(dev>=64 || !((configured.load()>>dev)&1ull)) {
		const cudaError_t e = cudaFuncSetAttribute(k_stream_collide_hyb<Q, COLL, ST, VF, ODD, SG, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if(e!=cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(stream_collide_hyb)");
		if(dev<64) configured.fetch_or(1ull<<dev);
	}
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#else
	sms = 2; per_sm = 1;
#endif
	const uint64_t all_blocks = (uint64_t)sms*(uint64_t)per_sm, blocks = all_blocks>2ull*(uint64_t)std::max(reserve, 0) ? all_blocks-(uint64_t)std::max(reserve, 0) : all_blocks, ntiles = (uint64_t)tiles_x*tiles_y*nz;
	if(ntiles==0ull) return FX3D_OK;
	if((uint64_t)tiles_x*tiles_y>0xFFFFFFFFull) { set_error("region has too many tile columns"); return FX3D_ERR_INVALID; }
	const dim3 grid((uint32_t)std::min<uint64_t>(ntiles, blocks), 1u, 1u);
	g_kind_launches[5]++;
	FX3D_LAUNCH_SMEM((k_stream_collide_hyb<Q, COLL, ST, VF, ODD, SG, MB>), grid, block, smem, stream, L, R, tiles_x, tiles_y);
	return check_launch("stream_collide (bulk loads, direct stores)");
}
template<int Q, int COLL, int ST, bool VF, bool SG = false, bool MB = false> static int launch_hyb(const Lattice& L, const Region& R, const dim3& block, void* stream, int reserve) {
	return L.odd ? launch_hyb_parity<Q, COLL, ST, VF, 1, SG, MB>(L, R, block, stream, reserve) : launch_hyb_parity<Q, COLL, ST, VF, 0, SG, MB>(L, R, block, stream, reserve);
}
template<int Q, int COLL, int ST, bool VF, bool SG = false, bool MB = false> static int launch_tma(const Lattice& L, const Region& R, void* stream, int reserve, const RowPeers& peers) {
	return L.odd ? launch_tma_parity<Q, COLL, ST, VF, 1, SG, MB>(L, R, stream, reserve, peers) : launch_tma_parity<Q, COLL, ST, VF, 0, SG, MB>(L, R, stream, reserve, peers);
}

template<int Q, int ST> int launch_stream_collide(const Lattice& L, const Region& R, int cells_per_thread, int collision, bool volume_force, void* stream, int reserve, int ext, const RowPeers* fused) {
	if(R.g1<=R.g0||R.y1<=R.y0||R.z1<=R.z0) return FX3D_OK;
	RowPeers peers; // the y/z neighbours whose rows this launch delivers itself; without them every row stays in this domain's memory
	for(int k=0; k<9; k++) peers.fi[k] = fused ? fused->fi[k] : nullptr;
	peers.fi[4] = L.fi;
	Lattice Lk = L; // the whole-row kernel routes halo rows by Hy/Hz: unfused launches see no y/z halos (rows then stay where the general rule puts them: here)
	if(!fused) { Lk.Hy = 0u; Lk.Hz = 0u; }
	if(cells_per_thread==-100) return tma_eligible<Q, ST>(L, R) ? FX3D_OK : 1; // query: would the whole-row kernel take this region?
#if defined(FX3D_TUNE_ONLY) // tuning builds (tools/build_variant.sh): only the whole-row kernel of the benchmark lines (SRT, no force, no extensions) is instantiated
	{
		if(cells_per_thread==32 && ext==0 && collision==COLL_SRT && !volume_force) { // the occupancy form, block shape from FX3D_OCC_BX
#ifndef FX3D_OCC_BX
#define FX3D_OCC_BX 128
#endif
			const dim3 block(FX3D_OCC_BX, 1u, 1u), grid((R.g1-R.g0+block.x-1u)/block.x, R.y1-R.y0, R.z1-R.z0);
			g_kind_launches[6]++;
			if(L.odd) FX3D_LAUNCH((k_stream_collide_occ<Q, COLL_SRT, ST, false, 1>), grid, block, stream, L, R); else FX3D_LAUNCH((k_stream_collide_occ<Q, COLL_SRT, ST, false, 0>), grid, block, stream, L, R);
			return check_launch("stream_collide (one cell per thread, high occupancy)");
		}
		const dim3 block = block_shape(R.g1-R.g0);
#if defined(FX3D_TUNE_TRT_VF) // ... or of the wind-tunnel line (TRT with VOLUME_FORCE)
		if(ext!=0 || cells_per_thread>0 || collision!=COLL_TRT || !volume_force || !tma_eligible<Q, ST>(L, R)) { set_error("tuning build: whole-row TRT+VOLUME_FORCE kernel only"); return FX3D_ERR_INVALID; }
		return launch_tma<Q, COLL_TRT, ST, true>(Lk, R, stream, reserve, peers);
#else
		if(ext!=0 || cells_per_thread>0 || collision!=COLL_SRT || volume_force || !tma_eligible<Q, ST>(L, R)) { set_error("tuning build: whole-row SRT kernel only"); return FX3D_ERR_INVALID; }
		return launch_tma<Q, COLL_SRT, ST, false>(Lk, R, stream, reserve, peers);
#endif
	}
#else
	if(ext!=0) { // SUBGRID (bit 0) and/or MOVING_BOUNDARIES (bit 1): the whole-row bulk-copy kernel (cells_per_thread 0, regions in groups of 4) or the general kernel (1)
		const dim3 block = block_shape(R.g1-R.g0);
		const bool sg = (ext&1)!=0, mb = (ext&2)!=0;
		if(cells_per_thread==0) {
			if(!tma_eligible<Q, ST>(L, R)) { // row segments: bulk loads + direct stores; else nothing launched, the caller falls back to the general kernel
				if(!tmaseg_eligible<Q, ST>(L, R, block)) return 1;
#define FX3D_HYB_EXT(COLL, VF) (sg ? (mb ? launch_hyb<Q, COLL, ST, VF, true, true>(L, R, block, stream, reserve) : launch_hyb<Q, COLL, ST, VF, true, false>(L, R, block, stream, reserve)) \
                                   : launch_hyb<Q, COLL, ST, VF, false, true>(L, R, block, stream, reserve))
				if(collision==COLL_SRT) return volume_force ? FX3D_HYB_EXT(COLL_SRT, true) : FX3D_HYB_EXT(COLL_SRT, false);
				return volume_force ? FX3D_HYB_EXT(COLL_TRT, true) : FX3D_HYB_EXT(COLL_TRT, false);
#undef FX3D_HYB_EXT
			}
#define FX3D_TMA_EXT(COLL, VF) (sg ? (mb ? launch_tma<Q, COLL, ST, VF, true, true>(Lk, R, stream, reserve, peers) : launch_tma<Q, COLL, ST, VF, true, false>(Lk, R, stream, reserve, peers)) \
                                   : launch_tma<Q, COLL, ST, VF, false, true>(Lk, R, stream, reserve, peers))
			if(collision==COLL_SRT) return volume_force ? FX3D_TMA_EXT(COLL_SRT, true) : FX3D_TMA_EXT(COLL_SRT, false);
			return volume_force ? FX3D_TMA_EXT(COLL_TRT, true) : FX3D_TMA_EXT(COLL_TRT, false);
#undef FX3D_TMA_EXT
		}
		const dim3 grid((R.g1-R.g0+block.x-1u)/block.x, (R.y1-R.y0+block.y-1u)/block.y, R.z1-R.z0);
		g_kind_launches[0]++;
#define FX3D_V1_EXT(COLL, VF) do { if(sg) FX3D_LAUNCH((k_stream_collide_v1<Q, COLL, ST, VF, true>), grid, block, stream, L, R); else FX3D_LAUNCH((k_stream_collide_v1<Q, COLL, ST, VF, false>), grid, block, stream, L, R); } while(0)
		if(collision==COLL_SRT) { if(volume_force) FX3D_V1_EXT(COLL_SRT, true); else FX3D_V1_EXT(COLL_SRT, false); }
		else { if(volume_force) FX3D_V1_EXT(COLL_TRT, true); else FX3D_V1_EXT(COLL_TRT, false); }
#undef FX3D_V1_EXT
		return check_launch("stream_collide (extensions)");
	}
	if(cells_per_thread<=0) { // persistent kernels; R.g0/g1 are in groups of pipe_cells<Q,ST>() cells. 0: bulk copies where the tile spans the row, else cp.async; -1: cp.async
		const dim3 block = block_shape(R.g1-R.g0);
		const bool rows = tma_eligible<Q, ST>(L, R);
		if(cells_per_thread==-2 && !rows) return 1; // -2: bulk copies of whole rows or nothing (regions in groups of 4 cells); 1 = "not eligible", no launch
		if(rows && (cells_per_thread==-2 || ((cells_per_thread==0 || cells_per_thread==-3) && pipe_cells<Q, ST>()==4))) {
			if(collision==COLL_SRT) return volume_force ? launch_tma<Q, COLL_SRT, ST, true>(Lk, R, stream, reserve, peers) : launch_tma<Q, COLL_SRT, ST, false>(Lk, R, stream, reserve, peers);
			return volume_force ? launch_tma<Q, COLL_TRT, ST, true>(Lk, R, stream, reserve, peers) : launch_tma<Q, COLL_TRT, ST, false>(Lk, R, stream, reserve, peers);
		}
		if(fused) return 1; // the caller asked for fused halo delivery, which only the whole-row kernel provides
		if((cells_per_thread==0 || cells_per_thread==-3) && pipe_cells<Q, ST>()==4 && tmaseg_eligible<Q, ST>(L, R, block)) { // row segments: bulk loads, direct stores
			if(collision==COLL_SRT) return volume_force ? launch_hyb<Q, COLL_SRT, ST, true>(L, R, block, stream, reserve) : launch_hyb<Q, COLL_SRT, ST, false>(L, R, block, stream, reserve);
			return volume_force ? launch_hyb<Q, COLL_TRT, ST, true>(L, R, block, stream, reserve) : launch_hyb<Q, COLL_TRT, ST, false>(L, R, block, stream, reserve);
		}
		if(collision==COLL_SRT) return volume_force ? launch_pipe<Q, COLL_SRT, ST, true>(L, R, block, stream, reserve) : launch_pipe<Q, COLL_SRT, ST, false>(L, R, block, stream, reserve);
		return volume_force ? launch_pipe<Q, COLL_TRT, ST, true>(L, R, block, stream, reserve) : launch_pipe<Q, COLL_TRT, ST, false>(L, R, block, stream, reserve);
	}
	const dim3 block = block_shape(R.g1-R.g0);
	const dim3 grid((R.g1-R.g0+block.x-1u)/block.x, (R.y1-R.y0+block.y-1u)/block.y, R.z1-R.z0);
	if(cells_per_thread==32) { // one cell per thread at high occupancy (R in cells)
		if((uint64_t)L.Nx*L.Ny*L.Nz>0xFFFFFFFFull) { set_error("the occupancy kernel indexes cells with 32 bits"); return FX3D_ERR_INVALID; }
		g_kind_launches[6]++;
#define FX3D_OCC(COLL, VF) do { if(L.odd) FX3D_LAUNCH((k_stream_collide_occ<Q, COLL, ST, VF, 1>), grid, block, stream, L, R); else FX3D_LAUNCH((k_stream_collide_occ<Q, COLL, ST, VF, 0>), grid, block, stream, L, R); } while(0)
		if(collision==COLL_SRT) { if(volume_force) FX3D_OCC(COLL_SRT, true); else FX3D_OCC(COLL_SRT, false); }
		else { if(volume_force) FX3D_OCC(COLL_TRT, true); else FX3D_OCC(COLL_TRT, false); }
#undef FX3D_OCC
		return check_launch("stream_collide (one cell per thread, high occupancy)");
	}
#define FX3D_SC(KERNEL, COLL, VF) FX3D_LAUNCH((KERNEL<Q, COLL, ST, VF>), grid, block, stream, L, R)
#define FX3D_SCV(K, COLL, VF) do { if(L.odd) FX3D_LAUNCH((k_stream_collide_vec<Q, COLL, ST, VF, K, 1>), grid, block, stream, L, R); else FX3D_LAUNCH((k_stream_collide_vec<Q, COLL, ST, VF, K, 0>), grid, block, stream, L, R); } while(0)
#define FX3D_SC_ALL(MACRO, ARG) \
	if(collision==COLL_SRT) { if(volume_force) MACRO(ARG, COLL_SRT, true); else MACRO(ARG, COLL_SRT, false); } \
	else                    { if(volume_force) MACRO(ARG, COLL_TRT, true); else MACRO(ARG, COLL_TRT, false); }
	g_kind_launches[cells_per_thread>1 ? 1 : 0]++;
	if(cells_per_thread==4) { FX3D_SC_ALL(FX3D_SCV, 4) }
	else if(cells_per_thread==2) { FX3D_SC_ALL(FX3D_SCV, 2) }
	else { FX3D_SC_ALL(FX3D_SC, k_stream_collide_v1) }
#undef FX3D_SC_ALL
#undef FX3D_SCV
#undef FX3D_SC
	return check_launch("stream_collide");
#endif
}
template int launch_stream_collide<FX3D_Q, FX3D_ST>(const Lattice&, const Region&, int, int, bool, void*, int, int, const RowPeers*);

} // namespace fx3d
