// sc_inst.cu -- stream_collide instantiations for one (velocity set, storage) pair; compiled once per pair with
// -DFX3D_Q=19|27 -DFX3D_ST=0|1|2 so that the six heavy translation units build in parallel.
#include "fx3d_internal.cuh"

namespace fx3d {

static inline dim3 block_shape(uint32_t nx) { // 128 threads; x extent = smallest power of two covering the row, capped at 128
	uint32_t bx = 1u;
	while(bx<nx && bx<128u) bx <<= 1;
	return dim3(bx, 128u/bx, 1u);
}

template<int Q, int ST> int launch_stream_collide(const Lattice& L, const Region& R, bool vector4, int collision, bool volume_force, void* stream) {
	if(R.g1<=R.g0||R.y1<=R.y0||R.z1<=R.z0) return FX3D_OK;
	const dim3 block = block_shape(R.g1-R.g0);
	const dim3 grid((R.g1-R.g0+block.x-1u)/block.x, (R.y1-R.y0+block.y-1u)/block.y, R.z1-R.z0);
#define FX3D_SC(KERNEL, COLL, VF) FX3D_LAUNCH((KERNEL<Q, COLL, ST, VF>), grid, block, stream, L, R)
	if(vector4) {
		if(collision==COLL_SRT) { if(volume_force) FX3D_SC(k_stream_collide_v4, COLL_SRT, true); else FX3D_SC(k_stream_collide_v4, COLL_SRT, false); }
		else                    { if(volume_force) FX3D_SC(k_stream_collide_v4, COLL_TRT, true); else FX3D_SC(k_stream_collide_v4, COLL_TRT, false); }
	} else {
		if(collision==COLL_SRT) { if(volume_force) FX3D_SC(k_stream_collide_v1, COLL_SRT, true); else FX3D_SC(k_stream_collide_v1, COLL_SRT, false); }
		else                    { if(volume_force) FX3D_SC(k_stream_collide_v1, COLL_TRT, true); else FX3D_SC(k_stream_collide_v1, COLL_TRT, false); }
	}
#undef FX3D_SC
	return check_launch("stream_collide");
}
template int launch_stream_collide<FX3D_Q, FX3D_ST>(const Lattice&, const Region&, bool, int, bool, void*);

} // namespace fx3d
