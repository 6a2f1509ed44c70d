// fx3d_lbm.cu -- C-ABI entry points for the LBM kernels (include/fx3d.h): argument checking, region decomposition,
// template dispatch. Replaces the Kernel objects LBM_Domain binds and launches (FluidX3D v3.7 src/lbm.cpp:127-129,
// 178-191, 1317-1354) and the launch glue of src/opencl.hpp:658-679.
#define FX3D_TU_LBM
#include "fx3d_internal.cuh"
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace fx3d {

static thread_local std::string t_error;
void set_error(const std::string& msg) { t_error = msg; }
std::atomic<uint64_t> g_launches{0ull};
std::atomic<int> g_variant{0};
std::atomic<int> g_interior_reserve{8};
std::atomic<uint64_t> g_kind_launches[8];

// Non-halo extent and the shell / interior split used to overlap the halo exchange with computation, in CELLS. The shell is
// every non-halo cell within one cell of a halo layer in y and z, and within XS cells in x, where XS (4, 2 or 1) is the largest
// cells-per-thread count that divides the non-halo row length: the split must not depend on which kernel (and which K) later
// runs a region, or the SHELL and INTERIOR passes of one step would overlap or leave cells out.
struct CellRegion { uint32_t x0, x1, y0, y1, z0, z1; };
static uint32_t x_shell_width(const Lattice& L) { const uint32_t inner = L.Nx-2u*L.Hx; return inner%4u==0u ? 4u : inner%2u==0u ? 2u : 1u; }
static void regions_of(const Lattice& L, int region, std::vector<CellRegion>& out) {
	const uint32_t x0 = L.Hx, x1 = L.Nx-L.Hx, y0 = L.Hy, y1 = L.Ny-L.Hy, z0 = L.Hz, z1 = L.Nz-L.Hz;
	if(region==FX3D_REGION_ALL || (L.Hx|L.Hy|L.Hz)==0u) {
		if(region!=FX3D_REGION_SHELL) out.push_back(CellRegion{ x0, x1, y0, y1, z0, z1 });
		return;
	}
	// interior box: shrink by one layer (XS cells along x) on every decomposed axis; empty if the domain is too thin
	const uint32_t xs = L.Hx ? x_shell_width(L) : 0u;
	const uint32_t ix0 = x0+xs, ix1 = x1>=x0+2u*xs ? x1-xs : x0+xs;
	const uint32_t iy0 = y0+L.Hy, iy1 = y1>=y0+2u*L.Hy ? y1-L.Hy : y0+L.Hy;
	const uint32_t iz0 = z0+L.Hz, iz1 = z1>=z0+2u*L.Hz ? z1-L.Hz : z0+L.Hz;
	const bool has_interior = ix1>ix0 && iy1>iy0 && iz1>iz0;
	if(region==FX3D_REGION_INTERIOR) {
		if(has_interior) out.push_back(CellRegion{ ix0, ix1, iy0, iy1, iz0, iz1 });
		return;
	}
	if(!has_interior) { out.push_back(CellRegion{ x0, x1, y0, y1, z0, z1 }); return; } // everything is shell
	if(L.Hz) { out.push_back(CellRegion{ x0, x1, y0, y1, z0, iz0 }); out.push_back(CellRegion{ x0, x1, y0, y1, iz1, z1 }); }
	if(L.Hy) { out.push_back(CellRegion{ x0, x1, y0, iy0, iz0, iz1 }); out.push_back(CellRegion{ x0, x1, iy1, y1, iz0, iz1 }); }
	if(L.Hx) { out.push_back(CellRegion{ x0, ix0, iy0, iy1, iz0, iz1 }); out.push_back(CellRegion{ ix1, x1, iy0, iy1, iz0, iz1 }); }
}
// the kernels address x in groups of K cells counted from the first non-halo cell (K>1) or in absolute cells (K==1)
static Region to_groups(const Lattice& L, const CellRegion& c, uint32_t K) {
	if(K<=1u) return Region{ c.x0, c.x1, c.y0, c.y1, c.z0, c.z1 };
	return Region{ (c.x0-L.Hx)/K, (c.x1-L.Hx)/K, c.y0, c.y1, c.z0, c.z1 };
}

static inline dim3 cell_block(uint32_t nx) { uint32_t bx = 1u; while(bx<nx && bx<128u) bx <<= 1; return dim3(bx, 128u/bx, 1u); }
static inline dim3 cell_grid(const Region& R, const dim3& b) { return dim3((R.g1-R.g0+b.x-1u)/b.x, (R.y1-R.y0+b.y-1u)/b.y, R.z1-R.z0); }
static inline Region all_cells(const Lattice& L) { return Region{ L.Hx, L.Nx-L.Hx, L.Hy, L.Ny-L.Hy, L.Hz, L.Nz-L.Hz }; }

#define FX3D_DISPATCH_Q_ST(Qv, STv, BODY) \
	if(Qv==19u) { if(STv==FX3D_FP32) { constexpr int Q = 19, ST = ST_FP32; BODY } else if(STv==FX3D_FP16S) { constexpr int Q = 19, ST = ST_FP16S; BODY } else { constexpr int Q = 19, ST = ST_FP16C; BODY } } \
	else        { if(STv==FX3D_FP32) { constexpr int Q = 27, ST = ST_FP32; BODY } else if(STv==FX3D_FP16S) { constexpr int Q = 27, ST = ST_FP16S; BODY } else { constexpr int Q = 27, ST = ST_FP16C; BODY } }

static uint32_t default_cells_per_thread(int storage) { (void)storage; return 4u; } // tuned on B200, see DESIGN.md
static bool default_pipelined(int region) { return region!=FX3D_REGION_SHELL; } // the persistent pipelined kernel marches in z; the one-cell shell slabs go to the vector kernel

// One region, one kernel. Kernel forms are tried in order of preference; a form that is not eligible for this region's shape
// launches nothing (launch_stream_collide returns 1) and the next one is tried, so a multi-slab SHELL can mix forms safely:
// every cell of the region list is advanced by exactly one launch.
static int stream_collide_region(const fx3d_lattice* lat, const Lattice& L, const CellRegion& c, int region, void* stream, const RowPeers* fused=nullptr, bool query=false) {
	const uint32_t width = c.x1-c.x0; // every x bound is a multiple of x_shell_width() cells from the first non-halo cell
	const int want = g_variant.load();
	const bool vf = (lat->features&FX3D_VOLUME_FORCE)!=0u;
	const int reserve = region==FX3D_REGION_INTERIOR ? g_interior_reserve.load() : 0;
	auto launch = [&](uint32_t K, int mode, int ext) -> int {
		const Region R = to_groups(L, c, K);
		int rc;
		FX3D_DISPATCH_Q_ST(lat->velocity_set, lat->storage, { rc = (launch_stream_collide<Q, ST>)(L, R, query ? -100 : mode, (int)lat->collision, vf, stream, reserve, ext, fused); })
		return rc;
	};
	const bool div4 = width%4u==0u && (c.x0-L.Hx)%4u==0u, div2 = width%2u==0u && (c.x0-L.Hx)%2u==0u;
	const bool cell_force = (lat->features&FX3D_FORCE_FIELD) && vf; // FORCE_FIELD only enters stream_collide through the volume-force terms (kernel.cpp:1494-1503,1550-1563)
	if(cell_force) { // per-cell force: the general kernel (one cell per thread) adds F[n] to (fx,fy,fz)
		if(fused || query) return 1;
		if(!L.F) { set_error("FORCE_FIELD lattice without a force field buffer (fx3d_lattice.F)"); return FX3D_ERR_INVALID; }
		const int ext = ((lat->features&FX3D_SUBGRID) ? 1 : 0)|((lat->features&FX3D_MOVING_BOUNDARIES) ? 2 : 0);
		return launch(1u, 1, ext ? ext : 4);
	}
	if(fused || query) { // fused y/z halo delivery exists in the whole-row bulk-copy kernel only (4 cells per thread)
		if(!div4 || want==1 || want==2 || want==4 || want==8 || want==32) return 1;
		const int ext = ((lat->features&FX3D_SUBGRID) ? 1 : 0)|((lat->features&FX3D_MOVING_BOUNDARIES) ? 2 : 0);
		return launch(4u, ext ? 0 : -2, ext);
	}
	if(lat->features&(FX3D_SUBGRID|FX3D_MOVING_BOUNDARIES)) { // widenings: bulk-copy / hybrid kernel where eligible, else the general kernel (any size)
		const int ext = ((lat->features&FX3D_SUBGRID) ? 1 : 0)|((lat->features&FX3D_MOVING_BOUNDARIES) ? 2 : 0);
		if(want!=1 && div4) { const int rc = launch(4u, 0, ext); if(rc!=1) return rc; }
		return launch(1u, 1, ext);
	}
	if(want==32) return launch(1u, 32, 0);
	const uint32_t pk = pipe_cells_of(lat->velocity_set, lat->storage);
	const bool pipelined = (want==8 || want==16 || (want==0 && default_pipelined(region))) && (pk==4u ? div4 : div2);
	if(pipelined) {
		if(want!=8 && pk==2u && div4) { const int rc = launch(4u, -2, 0); if(rc!=1) return rc; } // D3Q27 FP32: the bulk-copy kernel (4 cells per thread) where the tile spans whole rows
		return launch(pk, want==8 ? -1 : want==16 ? -3 : 0, 0);
	}
	uint32_t K = 1u;
	if(want!=1) {
		const uint32_t pref = want==2 ? 2u : want==4 ? 4u : default_cells_per_thread((int)lat->storage);
		if(pref==4u && div4) K = 4u; else if(div2) K = 2u;
	}
	return launch(K, (int)K, 0);
}

static int stream_collide_impl(const fx3d_lattice* lat, const Lattice& L, int region, void* stream) {
	std::vector<CellRegion> regs;
	regions_of(L, region, regs);
	for(const CellRegion& c : regs) {
		const int rc = stream_collide_region(lat, L, c, region, stream);
		if(rc==1) { set_error("stream_collide: no kernel form accepts this region"); return FX3D_ERR_INVALID; }
		if(rc!=FX3D_OK) return rc;
	}
	return FX3D_OK;
}

static int stream_collide_fused_impl(const fx3d_lattice* lat, const Lattice& L, void* const* fi_neighbours, void* stream, bool query) {
	std::vector<CellRegion> regs;
	regions_of(L, FX3D_REGION_ALL, regs);
	RowPeers peers;
	for(int k=0; k<9; k++) peers.fi[k] = fi_neighbours ? fi_neighbours[k] : nullptr;
	if(!query) for(int dz=-1; dz<=1; dz++) for(int dy=-1; dy<=1; dy++) { // every neighbour the kernel can route a row to must be there
		const bool needed = (dy==0 || L.Hy) && (dz==0 || L.Hz) && (dy!=0 || dz!=0);
		if(needed && !peers.fi[(dy+1)+3*(dz+1)]) { set_error("stream_collide_fused: a y/z neighbour's DDF buffer is null"); return FX3D_ERR_INVALID; }
	}
	const int rc = stream_collide_region(lat, L, regs[0], FX3D_REGION_ALL, stream, &peers, query);
	if(rc==1) { if(!query) set_error("stream_collide_fused: this lattice shape is not taken by the whole-row bulk-copy kernel (see fx3d_fused_halo_supported)"); return query ? 1 : FX3D_ERR_INVALID; }
	return rc;
}

} // namespace fx3d
using namespace fx3d;

extern "C" {

const char* fx3d_last_error(void) { return t_error.c_str(); }
int fx3d_set_kernel_variant(int variant) { if(variant!=0&&variant!=1&&variant!=2&&variant!=4&&variant!=8&&variant!=16&&variant!=32) { set_error("variant must be 0 (auto), 1 (general), 2 or 4 (cells per thread), 8 (pipelined, cp.async), 16 (pipelined, bulk copies where eligible), 32 (one cell per thread at high occupancy)"); return FX3D_ERR_INVALID; } g_variant = variant; return FX3D_OK; }
int fx3d_set_interior_reserve(int blocks) { if(blocks<0||blocks>1024) { set_error("reserve must be 0..1024 blocks"); return FX3D_ERR_INVALID; } g_interior_reserve = blocks; return FX3D_OK; }
int fx3d_stream_collide_launches(int kind, uint64_t* launches) { if(kind<0||kind>7||!launches) { set_error("kind must be 0..7"); return FX3D_ERR_INVALID; } *launches = g_kind_launches[kind].load(); return FX3D_OK; }
int fx3d_launch_count(uint64_t* launches) { if(!launches) return FX3D_ERR_INVALID; *launches = g_launches.load(); return FX3D_OK; }

size_t fx3d_fi_bytes(const fx3d_lattice* lat) {
	Lattice L;
	if(!make_lattice(lat, 0ull, 0.0f, 0.0f, 0.0f, L)) return 0u;
	return (size_t)(L.slot*lat->velocity_set*elem_bytes(lat->storage))+256u; // slack: the bulk-copy kernel reads one 16-byte chunk past a row segment
}
uint32_t fx3d_bytes_per_cell_per_step(const fx3d_lattice* lat) { // src/lbm.cpp:52-57
	if(!lat) return 0u;
	return lat->velocity_set*2u*(uint32_t)elem_bytes(lat->storage)+1u+((lat->features&FX3D_UPDATE_FIELDS) ? 16u : 0u)
		+((lat->features&FX3D_MOVING_BOUNDARIES) ? lat->velocity_set-1u : 0u); // the reference counts the neighbour flags every cell of a MOVING_BOUNDARIES build loads (:62-64)
}
float fx3d_relaxation_rate(float nu) {
	// def_w = to_string(1.0f/tau)+"f" (src/lbm.cpp:367): the reference formats 1/tau with 1+8 significant decimal digits in
	// float arithmetic (split_float, src/utilities.hpp:2599-2630) and the OpenCL compiler parses that literal back
	float x = 1.0f/(3.0f*nu+0.5f);
	if(!(x>0.0f) || x>3.4e38f) return x;
	int exponent = 0;
	if(x>=10.0f) {
		if(x>=1E32f) { x *= 1E-32f; exponent += 32; } if(x>=1E16f) { x *= 1E-16f; exponent += 16; } if(x>=1E8f) { x *= 1E-8f; exponent += 8; }
		if(x>=1E4f) { x *= 1E-4f; exponent += 4; } if(x>=1E2f) { x *= 1E-2f; exponent += 2; } if(x>=1E1f) { x *= 1E-1f; exponent += 1; }
	}
	if(x>0.0f && x<=1.0f) {
		if(x<1E-31f) { x *= 1E32f; exponent -= 32; } if(x<1E-15f) { x *= 1E16f; exponent -= 16; } if(x<1E-7f) { x *= 1E8f; exponent -= 8; }
		if(x<1E-3f) { x *= 1E4f; exponent -= 4; } if(x<1E-1f) { x *= 1E2f; exponent -= 2; } if(x<1E0f) { x *= 1E1f; exponent -= 1; }
	}
	uint32_t integral = (uint32_t)x;
	const float remainder = (x-(float)integral)*1E8f;
	uint32_t decimal = (uint32_t)remainder;
	if(remainder-(float)decimal>=0.5f) { decimal++; if(decimal>=100000000u) { decimal = 0u; integral++; if(integral>=10u) { integral = 1u; exponent++; } } }
	char s[48];
	std::snprintf(s, sizeof(s), "%u.%08uE%d", integral, decimal, exponent);
	return std::strtof(s, nullptr);
}

int fx3d_initialize(const fx3d_lattice* lat, fx3d_stream stream) {
	Lattice L;
	if(!make_lattice(lat, 1ull, 0.0f, 0.0f, 0.0f, L)) return FX3D_ERR_INVALID;
	if(int rc = use_device(lat->device)) return rc;
	const Region R = all_cells(L);
	const dim3 b = cell_block(R.g1-R.g0), g = cell_grid(R, b);
	FX3D_DISPATCH_Q_ST(lat->velocity_set, lat->storage, { FX3D_LAUNCH((k_initialize<Q, ST>), g, b, stream, L, R); })
	return check_launch("initialize");
}
int fx3d_stream_collide(const fx3d_lattice* lat, uint64_t t, float fx, float fy, float fz, int region, fx3d_stream stream) {
	Lattice L;
	if(!make_lattice(lat, t, fx, fy, fz, L)) return FX3D_ERR_INVALID;
	if(region<0||region>2) { set_error("invalid region"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(lat->device)) return rc;
	return stream_collide_impl(lat, L, region, stream);
}
int fx3d_stream_collide_fused(const fx3d_lattice* lat, uint64_t t, float fx, float fy, float fz, void* const* fi_neighbours, fx3d_stream stream) {
	Lattice L;
	if(!make_lattice(lat, t, fx, fy, fz, L)) return FX3D_ERR_INVALID;
	if(int rc = use_device(lat->device)) return rc;
	return stream_collide_fused_impl(lat, L, fi_neighbours, stream, false);
}
int fx3d_fused_halo_supported(const fx3d_lattice* lat) {
	Lattice L;
	if(!make_lattice(lat, 0ull, 0.0f, 0.0f, 0.0f, L)) return 0;
	if(!(L.Hy|L.Hz)) return 0; // nothing to fuse
	return stream_collide_fused_impl(lat, L, nullptr, nullptr, true)==FX3D_OK ? 1 : 0;
}
int fx3d_run_steps(const fx3d_lattice* lat, uint64_t t0, uint64_t steps, float fx, float fy, float fz, fx3d_stream stream) {
	Lattice L;
	if(!make_lattice(lat, t0, fx, fy, fz, L)) return FX3D_ERR_INVALID;
	if(lat->Dx*lat->Dy*lat->Dz!=1u) { set_error("fx3d_run_steps is for a single non-decomposed domain"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(lat->device)) return rc;
	for(uint64_t s=0ull; s<steps; s++) {
		L.odd = (uint32_t)((t0+s)&1ull);
		if(int rc = stream_collide_impl(lat, L, FX3D_REGION_ALL, stream)) return rc;
	}
	return FX3D_OK;
}
int fx3d_update_fields(const fx3d_lattice* lat, uint64_t t, float fx, float fy, float fz, fx3d_stream stream) {
	Lattice L;
	if(!make_lattice(lat, t, fx, fy, fz, L)) return FX3D_ERR_INVALID;
	if(int rc = use_device(lat->device)) return rc;
	const Region R = all_cells(L);
	const dim3 b = cell_block(R.g1-R.g0), g = cell_grid(R, b);
	if((lat->features&FX3D_FORCE_FIELD) && (lat->features&FX3D_VOLUME_FORCE) && !L.F) { set_error("FORCE_FIELD lattice without a force field buffer (fx3d_lattice.F)"); return FX3D_ERR_INVALID; }
	if(lat->features&FX3D_VOLUME_FORCE) { FX3D_DISPATCH_Q_ST(lat->velocity_set, lat->storage, { FX3D_LAUNCH((k_update_fields<Q, ST, true>), g, b, stream, L, R); }) }
	else { FX3D_DISPATCH_Q_ST(lat->velocity_set, lat->storage, { FX3D_LAUNCH((k_update_fields<Q, ST, false>), g, b, stream, L, R); }) }
	return check_launch("update_fields");
}

int fx3d_update_moving_boundaries(const fx3d_lattice* lat, fx3d_stream stream) {
	Lattice L;
	if(!make_lattice(lat, 0ull, 0.0f, 0.0f, 0.0f, L)) return FX3D_ERR_INVALID;
	if(!(lat->features&FX3D_MOVING_BOUNDARIES)) { set_error("update_moving_boundaries needs the MOVING_BOUNDARIES feature"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(lat->device)) return rc;
	const Region R = all_cells(L);
	const dim3 b = cell_block(R.g1-R.g0), g = cell_grid(R, b);
	if(lat->velocity_set==19u) FX3D_LAUNCH((k_update_moving_boundaries<19>), g, b, stream, L, R); else FX3D_LAUNCH((k_update_moving_boundaries<27>), g, b, stream, L, R);
	return check_launch("update_moving_boundaries");
}

// ---- FORCE_FIELD ----
static bool force_field_lattice(const fx3d_lattice* lat, uint64_t t, Lattice& L) {
	if(!make_lattice(lat, t, 0.0f, 0.0f, 0.0f, L)) return false;
	if(!(lat->features&FX3D_FORCE_FIELD)) { set_error("this call needs the FORCE_FIELD feature"); return false; }
	if(!L.F) { set_error("FORCE_FIELD lattice without a force field buffer (fx3d_lattice.F)"); return false; }
	return true;
}
int fx3d_update_force_field(const fx3d_lattice* lat, uint64_t t, fx3d_stream stream) {
	Lattice L;
	if(!force_field_lattice(lat, t, L)) return FX3D_ERR_INVALID;
	if(int rc = use_device(lat->device)) return rc;
	const Region R = all_cells(L);
	const dim3 b = cell_block(R.g1-R.g0), g = cell_grid(R, b);
	FX3D_DISPATCH_Q_ST(lat->velocity_set, lat->storage, { FX3D_LAUNCH((k_update_force_field<Q, ST>), g, b, stream, L, R); })
	return check_launch("update_force_field");
}
int fx3d_reset_force_field(const fx3d_lattice* lat, fx3d_stream stream) { // kernel reset_force_field, kernel.cpp:1885-1889: halo included
	Lattice L;
	if(!force_field_lattice(lat, 0ull, L)) return FX3D_ERR_INVALID;
	return fx3d_memset(lat->device, L.F, 0, (size_t)(3ull*(uint64_t)L.Nx*L.Ny*L.Nz*4ull), stream);
}
size_t fx3d_object_scratch_bytes(const fx3d_lattice* lat) {
	if(!lat) return 0u;
	const uint64_t N = (uint64_t)lat->Nx*lat->Ny*lat->Nz;
	return (size_t)(((N+OBJECT_GROUP-1u)/OBJECT_GROUP)*sizeof(ObjectPartial));
}
static int object_sum_impl(const fx3d_lattice* lat, uint32_t kind, uint8_t flag_marker, float cx, float cy, float cz, float* object_sum, void* scratch, fx3d_stream stream) {
	Lattice L;
	if(!force_field_lattice(lat, 0ull, L)) return FX3D_ERR_INVALID;
	if(!object_sum || !scratch) { set_error("object_sum / scratch buffer is null"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(lat->device)) return rc;
	const uint64_t groups = ((uint64_t)L.Nx*L.Ny*L.Nz+OBJECT_GROUP-1u)/OBJECT_GROUP;
	if(groups>0x7FFFFFFFull) { set_error("lattice too large for the object sums"); return FX3D_ERR_INVALID; }
	ObjectPartial* partial = reinterpret_cast<ObjectPartial*>(scratch);
	const dim3 g((uint32_t)groups), b(OBJECT_GROUP);
	const uint32_t smem = 4u*OBJECT_GROUP*4u;
	if(kind==0u) FX3D_LAUNCH_SMEM((k_object_partial<0>), g, b, smem, stream, L, flag_marker, cx, cy, cz, partial);
	else if(kind==1u) FX3D_LAUNCH_SMEM((k_object_partial<1>), g, b, smem, stream, L, flag_marker, cx, cy, cz, partial);
	else FX3D_LAUNCH_SMEM((k_object_partial<2>), g, b, smem, stream, L, flag_marker, cx, cy, cz, partial);
	FX3D_LAUNCH(k_object_total, dim3(1u), dim3(32u), stream, partial, (uint32_t)groups, kind, object_sum);
	return check_launch("object sum");
}
int fx3d_object_center_of_mass(const fx3d_lattice* lat, uint8_t flag_marker, float* object_sum, void* scratch, fx3d_stream stream) { return object_sum_impl(lat, 0u, flag_marker, 0.0f, 0.0f, 0.0f, object_sum, scratch, stream); }
int fx3d_object_force(const fx3d_lattice* lat, uint8_t flag_marker, float* object_sum, void* scratch, fx3d_stream stream) { return object_sum_impl(lat, 1u, flag_marker, 0.0f, 0.0f, 0.0f, object_sum, scratch, stream); }
int fx3d_object_torque(const fx3d_lattice* lat, uint8_t flag_marker, float cx, float cy, float cz, float* object_sum, void* scratch, fx3d_stream stream) { return object_sum_impl(lat, 2u, flag_marker, cx, cy, cz, object_sum, scratch, stream); }

int fx3d_voxelize_mesh(const fx3d_lattice* lat, int Ox, int Oy, int Oz, uint32_t direction, uint64_t t, uint8_t flag, const float* p0, const float* p1, const float* p2, const float* bbu, fx3d_stream stream) {
	Lattice L;
	if(!make_lattice(lat, t, 0.0f, 0.0f, 0.0f, L)) return FX3D_ERR_INVALID;
	if(direction>2u || !p0 || !p1 || !p2 || !bbu) { set_error("voxelize_mesh: direction must be 0, 1 or 2 and the triangle buffers must not be null"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(lat->device)) return rc;
	VoxelizeArgs V;
	std::memcpy(&V.triangle_number, &bbu[0], 4);
	V.x0 = bbu[1]; V.y0 = bbu[2]; V.z0 = bbu[3]; V.x1 = bbu[4]; V.y1 = bbu[5]; V.z1 = bbu[6]; V.cx = bbu[7]; V.cy = bbu[8]; V.cz = bbu[9];
	V.ux = bbu[10]; V.uy = bbu[11]; V.uz = bbu[12]; V.rx = bbu[13]; V.ry = bbu[14]; V.rz = bbu[15]; V.Ox = Ox; V.Oy = Oy; V.Oz = Oz;
	const uint64_t A = direction==0u ? (uint64_t)L.Ny*L.Nz : direction==1u ? (uint64_t)L.Nz*L.Nx : (uint64_t)L.Nx*L.Ny;
	if(A>0xFFFFFFFFull) { set_error("face too large"); return FX3D_ERR_INVALID; }
	const dim3 b(128u), g((uint32_t)((A+127ull)/128ull));
	const uint32_t smem = VOX_CHUNK*9u*4u, t_odd = (uint32_t)(t&1ull);
	FX3D_DISPATCH_Q_ST(lat->velocity_set, lat->storage, { FX3D_LAUNCH_SMEM((k_voxelize_mesh<Q, ST>), g, b, smem, stream, L, direction, t_odd, flag, p0, p1, p2, V); })
	return check_launch("voxelize_mesh");
}
int fx3d_unvoxelize_mesh(const fx3d_lattice* lat, int Ox, int Oy, int Oz, uint8_t flag, float x0, float y0, float z0, float x1, float y1, float z1, fx3d_stream stream) {
	Lattice L;
	if(!make_lattice(lat, 0ull, 0.0f, 0.0f, 0.0f, L)) return FX3D_ERR_INVALID;
	if(int rc = use_device(lat->device)) return rc;
	const uint64_t N = (uint64_t)L.Nx*L.Ny*L.Nz;
	FX3D_LAUNCH(k_unvoxelize_mesh, dim3((uint32_t)((N+127ull)/128ull)), dim3(128u), stream, L, flag, x0, y0, z0, x1, y1, z1, Ox, Oy, Oz);
	return check_launch("unvoxelize_mesh");
}

static bool face_setup(const fx3d_lattice* lat, uint32_t axis, uint64_t t, Lattice& L, dim3& g, dim3& b) {
	if(!make_lattice(lat, t, 0.0f, 0.0f, 0.0f, L)) return false;
	if(axis>2u) { set_error("axis must be 0, 1 or 2"); return false; }
	if((axis==0u&&L.Nx<3u)||(axis==1u&&L.Ny<3u)||(axis==2u&&L.Nz<3u)) { set_error("axis too thin for a halo transfer"); return false; }
	const uint64_t A = axis==0u ? (uint64_t)L.Ny*L.Nz : axis==1u ? (uint64_t)L.Nz*L.Nx : (uint64_t)L.Nx*L.Ny;
	if(A>0xFFFFFFFFull) { set_error("face too large"); return false; }
	b = dim3(128u, 1u, 1u); g = dim3((uint32_t)((A+127ull)/128ull), 1u, 1u);
	return true;
}
size_t fx3d_transfer_bytes(const fx3d_lattice* lat) { // src/lbm.cpp:1309-1315
	if(!lat) return 0u;
	uint64_t Amax = 0ull;
	if(lat->Dx>1u) Amax = std::max(Amax, (uint64_t)lat->Ny*lat->Nz);
	if(lat->Dy>1u) Amax = std::max(Amax, (uint64_t)lat->Nz*lat->Nx);
	if(lat->Dz>1u) Amax = std::max(Amax, (uint64_t)lat->Nx*lat->Ny);
	const uint64_t per = std::max<uint64_t>((lat->velocity_set==19u ? 5u : 9u)*elem_bytes(lat->storage), 17u);
	return (size_t)(Amax*per);
}
int fx3d_transfer_extract_fi(const fx3d_lattice* lat, uint32_t axis, uint64_t t, void* bp, void* bm, fx3d_stream stream) {
	Lattice L; dim3 g, b;
	if(!face_setup(lat, axis, t, L, g, b)) return FX3D_ERR_INVALID;
	if(int rc = use_device(lat->device)) return rc;
	FX3D_DISPATCH_Q_ST(lat->velocity_set, lat->storage, { g.y = (uint32_t)transfers<Q>(); FX3D_LAUNCH((k_transfer_fi<Q, ST, true>), g, b, stream, L, axis, bp, bm); })
	return check_launch("transfer_extract_fi");
}
int fx3d_transfer_insert_fi(const fx3d_lattice* lat, uint32_t axis, uint64_t t, const void* bp, const void* bm, fx3d_stream stream) {
	Lattice L; dim3 g, b;
	if(!face_setup(lat, axis, t, L, g, b)) return FX3D_ERR_INVALID;
	if(int rc = use_device(lat->device)) return rc;
	FX3D_DISPATCH_Q_ST(lat->velocity_set, lat->storage, { g.y = (uint32_t)transfers<Q>(); FX3D_LAUNCH((k_transfer_fi<Q, ST, false>), g, b, stream, L, axis, const_cast<void*>(bp), const_cast<void*>(bm)); })
	return check_launch("transfer_insert_fi");
}
int fx3d_transfer_extract_rho_u_flags(const fx3d_lattice* lat, uint32_t axis, uint64_t t, void* bp, void* bm, fx3d_stream stream) {
	Lattice L; dim3 g, b;
	if(!face_setup(lat, axis, t, L, g, b)) return FX3D_ERR_INVALID;
	if(int rc = use_device(lat->device)) return rc;
	FX3D_LAUNCH((k_transfer_rho_u_flags<true>), g, b, stream, L, axis, bp, bm);
	return check_launch("transfer_extract_rho_u_flags");
}
int fx3d_transfer_insert_rho_u_flags(const fx3d_lattice* lat, uint32_t axis, uint64_t t, const void* bp, const void* bm, fx3d_stream stream) {
	Lattice L; dim3 g, b;
	if(!face_setup(lat, axis, t, L, g, b)) return FX3D_ERR_INVALID;
	if(int rc = use_device(lat->device)) return rc;
	FX3D_LAUNCH((k_transfer_rho_u_flags<false>), g, b, stream, L, axis, const_cast<void*>(bp), const_cast<void*>(bm));
	return check_launch("transfer_insert_rho_u_flags");
}
int fx3d_transfer_extract_flags(const fx3d_lattice* lat, uint32_t axis, uint64_t t, void* bp, void* bm, fx3d_stream stream) {
	Lattice L; dim3 g, b;
	if(!face_setup(lat, axis, t, L, g, b)) return FX3D_ERR_INVALID;
	if(!bp||!bm) { set_error("transfer buffers are null"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(lat->device)) return rc;
	FX3D_LAUNCH((k_transfer_flags<true>), g, b, stream, L, axis, reinterpret_cast<uint8_t*>(bp), reinterpret_cast<uint8_t*>(bm));
	return check_launch("transfer_extract_flags");
}
int fx3d_transfer_insert_flags(const fx3d_lattice* lat, uint32_t axis, uint64_t t, const void* bp, const void* bm, fx3d_stream stream) {
	Lattice L; dim3 g, b;
	if(!face_setup(lat, axis, t, L, g, b)) return FX3D_ERR_INVALID;
	if(!bp||!bm) { set_error("transfer buffers are null"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(lat->device)) return rc;
	FX3D_LAUNCH((k_transfer_flags<false>), g, b, stream, L, axis, reinterpret_cast<uint8_t*>(const_cast<void*>(bp)), reinterpret_cast<uint8_t*>(const_cast<void*>(bm)));
	return check_launch("transfer_insert_flags");
}
int fx3d_transfer_extract_F(const fx3d_lattice* lat, uint32_t axis, uint64_t t, void* bp, void* bm, fx3d_stream stream) {
	Lattice L; dim3 g, b;
	if(!face_setup(lat, axis, t, L, g, b)) return FX3D_ERR_INVALID;
	if(!L.F||!bp||!bm) { set_error("transfer_extract_F needs a FORCE_FIELD lattice and two buffers"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(lat->device)) return rc;
	FX3D_LAUNCH((k_transfer_F<true>), g, b, stream, L, axis, reinterpret_cast<float*>(bp), reinterpret_cast<float*>(bm));
	return check_launch("transfer_extract_F");
}
int fx3d_transfer_insert_F(const fx3d_lattice* lat, uint32_t axis, uint64_t t, const void* bp, const void* bm, fx3d_stream stream) {
	Lattice L; dim3 g, b;
	if(!face_setup(lat, axis, t, L, g, b)) return FX3D_ERR_INVALID;
	if(!L.F||!bp||!bm) { set_error("transfer_insert_F needs a FORCE_FIELD lattice and two buffers"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(lat->device)) return rc;
	FX3D_LAUNCH((k_transfer_F<false>), g, b, stream, L, axis, reinterpret_cast<float*>(const_cast<void*>(bp)), reinterpret_cast<float*>(const_cast<void*>(bm)));
	return check_launch("transfer_insert_F");
}
int fx3d_exchange_flags(const fx3d_lattice* lat, uint32_t axis, const uint8_t* flags_plus, const uint8_t* flags_minus, fx3d_stream stream) {
	Lattice L; dim3 g, b;
	if(!face_setup(lat, axis, 0ull, L, g, b)) return FX3D_ERR_INVALID;
	if(!flags_plus||!flags_minus) { set_error("neighbour flag buffers are null"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(lat->device)) return rc;
	FX3D_LAUNCH(k_exchange_flags, g, b, stream, L, axis, flags_plus, flags_minus);
	return check_launch("exchange_flags");
}
int fx3d_exchange_F(const fx3d_lattice* lat, uint32_t axis, const float* F_plus, const float* F_minus, fx3d_stream stream) {
	Lattice L; dim3 g, b;
	if(!face_setup(lat, axis, 0ull, L, g, b)) return FX3D_ERR_INVALID;
	if(!L.F||!F_plus||!F_minus) { set_error("exchange_F needs a FORCE_FIELD lattice and the neighbours' force fields"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(lat->device)) return rc;
	FX3D_LAUNCH(k_exchange_F, g, b, stream, L, axis, F_plus, F_minus);
	return check_launch("exchange_F");
}
int fx3d_exchange_fi(const fx3d_lattice* lat, uint32_t axis, uint64_t t, const void* fi_plus, const void* fi_minus, fx3d_stream stream) {
	Lattice L; dim3 g, b;
	if(!face_setup(lat, axis, t, L, g, b)) return FX3D_ERR_INVALID;
	if(!fi_plus||!fi_minus) { set_error("neighbour DDF buffers are null"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(lat->device)) return rc;
	if(axis==0u) { // x faces are one element per row
		FX3D_DISPATCH_Q_ST(lat->velocity_set, lat->storage, { g.y = (uint32_t)transfers<Q>(); FX3D_LAUNCH((k_exchange_fi<Q, ST>), g, b, stream, L, axis, fi_plus, fi_minus); })
	} else { // y and z faces are whole x-rows: vector copies
		const uint32_t rows = axis==1u ? L.Nz : L.Ny, nvec = L.px*(uint32_t)elem_bytes(lat->storage)/16u;
		b = dim3(nvec>=128u ? 128u : nvec>=64u ? 64u : 32u, 1u, 1u);
		FX3D_DISPATCH_Q_ST(lat->velocity_set, lat->storage, { g = dim3(rows, (uint32_t)transfers<Q>(), 2u); FX3D_LAUNCH((k_exchange_fi_rows<Q, ST>), g, b, stream, L, axis, fi_plus, fi_minus); })
	}
	return check_launch("exchange_fi");
}
int fx3d_exchange_rho_u_flags(const fx3d_lattice* lat, uint32_t axis, const float* rho_plus, const float* u_plus, const uint8_t* flags_plus,
	const float* rho_minus, const float* u_minus, const uint8_t* flags_minus, fx3d_stream stream) {
	Lattice L; dim3 g, b;
	if(!face_setup(lat, axis, 0ull, L, g, b)) return FX3D_ERR_INVALID;
	if(!rho_plus||!u_plus||!flags_plus||!rho_minus||!u_minus||!flags_minus) { set_error("neighbour field buffers are null"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(lat->device)) return rc;
	const PeerFields plus{ rho_plus, u_plus, flags_plus }, minus{ rho_minus, u_minus, flags_minus };
	FX3D_LAUNCH(k_exchange_rho_u_flags, g, b, stream, L, axis, plus, minus);
	return check_launch("exchange_rho_u_flags");
}

int fx3d_codec_encode(int device, int storage, const float* in, uint16_t* out, size_t count, fx3d_stream stream) {
	if(storage!=FX3D_FP16S&&storage!=FX3D_FP16C) { set_error("codec entry points are for FP16S / FP16C"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(device)) return rc;
	const dim3 b(128u), g((uint32_t)((count+127u)/128u));
	if(count==0u) return FX3D_OK;
	if(storage==FX3D_FP16S) FX3D_LAUNCH((k_codec_encode<ST_FP16S>), g, b, stream, in, out, (uint64_t)count);
	else FX3D_LAUNCH((k_codec_encode<ST_FP16C>), g, b, stream, in, out, (uint64_t)count);
	return check_launch("codec_encode");
}
int fx3d_codec_decode(int device, int storage, const uint16_t* in, float* out, size_t count, fx3d_stream stream) {
	if(storage!=FX3D_FP16S&&storage!=FX3D_FP16C) { set_error("codec entry points are for FP16S / FP16C"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(device)) return rc;
	const dim3 b(128u), g((uint32_t)((count+127u)/128u));
	if(count==0u) return FX3D_OK;
	if(storage==FX3D_FP16S) FX3D_LAUNCH((k_codec_decode<ST_FP16S>), g, b, stream, in, out, (uint64_t)count);
	else FX3D_LAUNCH((k_codec_decode<ST_FP16C>), g, b, stream, in, out, (uint64_t)count);
	return check_launch("codec_decode");
}

} // extern "C"
