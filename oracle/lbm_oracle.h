/*
 * lbm_oracle.h -- CPU oracle for the FluidX3D lattice-Boltzmann hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker or the reported CPU baseline. The product path
 * (fluidx3d_b200/, include/fx3d.h) never links or calls it.
 *
 * It is a plain-C restatement (run-time parametrised, table driven) of the reference's OpenCL C
 * device code and the host sequencing around it; every function cites the reference file:line it
 * follows (paths relative to the FluidX3D v3.7 source tree).
 *
 * Parity pin: this restatement is checked bit-for-bit against (a) oracle/_ref, the reference's own
 * kernel.cpp source compiled natively through oracle/ref/ocl_shim.hpp (built only where the
 * reference tree is mounted), and (b) the golden vectors under tests/golden/ that were generated
 * from (a) by oracle/make_golden.py.
 */
#ifndef LBM_ORACLE_H
#define LBM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_FP32 = 0, ORC_FP16S = 1, ORC_FP16C = 2 };
enum { ORC_SRT = 0, ORC_TRT = 1 };
enum { ORC_VOLUME_FORCE = 1u, ORC_EQUILIBRIUM_BOUNDARIES = 2u, ORC_UPDATE_FIELDS = 4u, ORC_SUBGRID = 8u, ORC_MOVING_BOUNDARIES = 16u, ORC_FORCE_FIELD = 32u };

/* One LBM_Domain worth of compile-time constants of the reference (src/lbm.cpp:334-425), made run-time. */
typedef struct orc_grid {
	uint32_t Nx, Ny, Nz;   /* local lattice size, halo layers included (def_Nx..def_Nz) */
	uint32_t Dx, Dy, Dz;   /* number of domains per axis; only gates is_halo() (def_Dx..def_Dz) */
	uint32_t Q;            /* velocity set: 19 or 27 */
	uint32_t collision;    /* ORC_SRT | ORC_TRT */
	uint32_t storage;      /* ORC_FP32 | ORC_FP16S | ORC_FP16C */
	uint32_t features;     /* ORC_VOLUME_FORCE | ORC_EQUILIBRIUM_BOUNDARIES | ORC_UPDATE_FIELDS | ORC_SUBGRID | ORC_MOVING_BOUNDARIES */
	float w;               /* relaxation rate def_w = 1/tau, as the device sees it */
} orc_grid;

/* storage codecs (src/lbm.cpp:410-425, src/kernel.cpp:848-859) */
uint16_t orc_fp16s_encode(float x);
float    orc_fp16s_decode(uint16_t h);
uint16_t orc_fp16c_encode(float x);
float    orc_fp16c_decode(uint16_t h);

/* decimal round trip of def_w (src/utilities.hpp:2599-2630,2745-2754; src/lbm.cpp:367) */
float orc_w_from_nu(float nu);
int   orc_float_to_string(float x, char* out, int cap);

/* kernels; fi is float[Q*N] for FP32, uint16_t[Q*N] otherwise; u is float[3*N] SoA */
void orc_initialize(const orc_grid* g, void* fi, const float* rho, float* u, uint8_t* flags);
void orc_stream_collide(const orc_grid* g, void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t, float fx, float fy, float fz);
void orc_update_fields(const orc_grid* g, const void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t, float fx, float fy, float fz);
/* FORCE_FIELD: the same two kernels with the per-cell force F (float[3N] SoA) added to (fx,fy,fz), src/kernel.cpp:1497-1503,1821-1827 */
void orc_stream_collide_F(const orc_grid* g, void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t, float fx, float fy, float fz, const float* F);
void orc_update_fields_F(const orc_grid* g, const void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t, float fx, float fy, float fz, const float* F);
void orc_update_force_field(const orc_grid* g, const void* fi, const uint8_t* flags, uint64_t t, float* F); /* src/kernel.cpp:1873-1884 */
void orc_reset_force_field(const orc_grid* g, float* F);                                                     /* :1885-1889 */
/* object_center_of_mass (kind 0) / object_force (1) / object_torque (2), :1901-1959, with the reduction order fixed (see lbm_oracle.c); out4 = x,y,z,count bits */
void orc_object_sum(const orc_grid* g, uint32_t kind, const float* F, const uint8_t* flags, uint8_t flag_marker, float cx, float cy, float cz, uint32_t group, float* out4);
void orc_transfer_extract_F(const orc_grid* g, uint32_t axis, uint64_t t, void* buf_p, void* buf_m, const float* F); /* :2173-2196 */
void orc_transfer_insert_F(const orc_grid* g, uint32_t axis, uint64_t t, const void* buf_p, const void* buf_m, float* F);
/* MOVING_BOUNDARIES: mark / unmark the cells next to TYPE_S cells with non-zero velocity as TYPE_MS (src/kernel.cpp:1432-1450) */
void orc_update_moving_boundaries(const orc_grid* g, const float* u, uint8_t* flags);

/* halo transfer kernels; axis 0|1|2; buffers hold transfers*A elements of the storage type (fi) or 17*A bytes */
uint32_t orc_transfers(const orc_grid* g);
uint64_t orc_area(const orc_grid* g, uint32_t axis);
void orc_transfer_extract_fi(const orc_grid* g, uint32_t axis, uint64_t t, void* buf_p, void* buf_m, const void* fi);
void orc_transfer_insert_fi(const orc_grid* g, uint32_t axis, uint64_t t, const void* buf_p, const void* buf_m, void* fi);
void orc_transfer_extract_rho_u_flags(const orc_grid* g, uint32_t axis, uint64_t t, void* buf_p, void* buf_m, const float* rho, const float* u, const uint8_t* flags);
void orc_transfer_insert_rho_u_flags(const orc_grid* g, uint32_t axis, uint64_t t, const void* buf_p, const void* buf_m, float* rho, float* u, uint8_t* flags);

/* GPU voxeliser: src/kernel.cpp:2267-2357; Ox,Oy,Oz = offset of the domain in the global grid; p0,p1,p2 = 3 floats per
 * triangle; bbu = the 16-float parameter block of LBM_Domain::voxelize_mesh_on_device (src/lbm.cpp:279-296) */
void orc_voxelize_mesh(const orc_grid* g, int Ox, int Oy, int Oz, uint32_t direction, void* fi, float* u, uint8_t* flags, uint64_t t, uint8_t flag,
	const float* p0, const float* p1, const float* p2, const float* bbu);
void orc_unvoxelize_mesh(const orc_grid* g, int Ox, int Oy, int Oz, uint8_t* flags, uint8_t flag, float x0, float y0, float z0, float x1, float y1, float z1);

void orc_set_threads(int n); /* OpenMP threads used by the kernels above (0 = library default) */
int  orc_get_threads(void);

#ifdef __cplusplus
}
#endif
#endif
