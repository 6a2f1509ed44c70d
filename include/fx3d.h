/*
 * fx3d.h -- C ABI of libfx3d_cuda.so, the B200 (sm_100a) device layer behind the FluidX3D host API.
 *
 * This is the drop-in boundary: it replaces the reference's OpenCL wrapper src/opencl.hpp (Device_Info :89-198,
 * get_devices :222-252, Device :284-340, Memory<T> :342-614, Kernel :616-692) and the OpenCL C kernels the LBM
 * classes launch through it (src/kernel.cpp), as called from src/lbm.cpp. Plain pointers and sizes only; every
 * function returns 0 on success or a negative fx3d_status, never exits, and leaves a message for
 * fx3d_last_error(). All launches are asynchronous on the given stream unless stated otherwise. There is no CPU
 * path: without a CUDA device every call fails with FX3D_ERR_NO_DEVICE.
 *
 * Host-side C++ wrappers that keep the reference's class and method names live in fluidx3d_b200/host/.
 */
#ifndef FX3D_H
#define FX3D_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum fx3d_status {
	FX3D_OK = 0,
	FX3D_ERR_NO_DEVICE = -1,   /* no CUDA device / driver (reference: "No OpenCL devices are available", opencl.hpp:238-245) */
	FX3D_ERR_INVALID = -2,     /* invalid argument or unsupported combination */
	FX3D_ERR_OUT_OF_MEMORY = -3, /* reference: "Memory size is too large", opencl.hpp:386 */
	FX3D_ERR_CUDA = -4,        /* any other CUDA runtime error, text in fx3d_last_error() */
	FX3D_ERR_TIMEOUT = -5      /* a peer did not reach the halo rendezvous in time */
} fx3d_status;

enum { FX3D_FP32 = 0, FX3D_FP16S = 1, FX3D_FP16C = 2 };                 /* defines.hpp: (none) | FP16S | FP16C */
enum { FX3D_SRT = 0, FX3D_TRT = 1 };                                    /* defines.hpp: SRT | TRT */
enum { FX3D_VOLUME_FORCE = 1, FX3D_EQUILIBRIUM_BOUNDARIES = 2, FX3D_UPDATE_FIELDS = 4, FX3D_SUBGRID = 8, FX3D_MOVING_BOUNDARIES = 16, FX3D_FORCE_FIELD = 32 }; /* defines.hpp extension flags on the path;
 * SUBGRID (Smagorinsky-Lilly, kernel.cpp:1579-1593), MOVING_BOUNDARIES (kernel.cpp:1104-1113,1378-1387,1432-1450) and FORCE_FIELD (kernel.cpp:1497-1503,
 * 1873-1959) are the widenings beyond the north_star feature set */
enum { FX3D_REGION_ALL = 0, FX3D_REGION_SHELL = 1, FX3D_REGION_INTERIOR = 2 };

const char* fx3d_last_error(void); /* thread-local, valid until the next failing call on this thread */

/* ---- devices: replaces Device_Info / get_devices (opencl.hpp:89-252) ---- */
typedef struct fx3d_device_info {
	char name[256];
	int id;                    /* CUDA ordinal */
	int cc_major, cc_minor;
	int sm_count;              /* Device_Info::compute_units */
	int clock_mhz;             /* Device_Info::clock_frequency */
	uint64_t memory_bytes;     /* Device_Info::memory (the reference keeps MB) */
	uint64_t l2_bytes;
	float tflops_fp32;         /* Device_Info::tflops */
} fx3d_device_info;
int fx3d_device_count(int* count);
int fx3d_device_get_info(int device, fx3d_device_info* info);
int fx3d_device_enable_peer(int device, int peer); /* both directions are the caller's job; no-op if device==peer */
int fx3d_device_sync(int device);                  /* Device::finish_queue for every stream of the device */

/* ---- streams and events: replaces the per-device in-order cl::CommandQueue (opencl.hpp:311) ---- */
typedef struct fx3d_stream_s* fx3d_stream; /* a cudaStream_t */
typedef struct fx3d_event_s* fx3d_event;   /* a cudaEvent_t */
int fx3d_stream_create(int device, fx3d_stream* stream);
int fx3d_stream_destroy(int device, fx3d_stream stream);
int fx3d_stream_sync(int device, fx3d_stream stream);                 /* Device::finish_queue (opencl.hpp:335) */
int fx3d_event_create(int device, fx3d_event* event);
int fx3d_event_destroy(int device, fx3d_event event);
int fx3d_event_record(int device, fx3d_event event, fx3d_stream stream);
int fx3d_event_sync(int device, fx3d_event event);
int fx3d_event_elapsed_ms(fx3d_event start, fx3d_event stop, float* ms);
int fx3d_stream_wait_event(int device, fx3d_stream stream, fx3d_event event);

/* ---- memory: replaces Memory<T> allocation and transfers (opencl.hpp:361-388, 510-531, 608-611) ---- */
int fx3d_malloc(int device, size_t bytes, void** ptr);                /* device buffer, zero-filled */
int fx3d_free(int device, void* ptr);
int fx3d_host_alloc(size_t bytes, void** ptr);                        /* page-locked host buffer */
int fx3d_host_free(void* ptr);
int fx3d_memcpy_h2d(int device, void* dst, const void* src, size_t bytes, fx3d_stream stream, int blocking); /* write_to_device */
int fx3d_memcpy_d2h(int device, void* dst, const void* src, size_t bytes, fx3d_stream stream, int blocking); /* read_from_device */
int fx3d_memset(int device, void* dst, int value, size_t bytes, fx3d_stream stream);
int fx3d_fill_f32(int device, float* dst, float value, size_t count, fx3d_stream stream);                    /* Memory::reset(value) on device */

/* cross-process sharing of device buffers (one process per GPU): cudaIpc handles are 64 opaque bytes */
int fx3d_ipc_get_handle(int device, void* ptr, void* handle64);
int fx3d_ipc_open_handle(int device, const void* handle64, void** ptr);
int fx3d_ipc_close_handle(int device, void* ptr);

/* ---- one LBM_Domain as the device sees it (constants LBM_Domain::device_defines bakes into the JIT source,
 * src/lbm.cpp:334-425, plus the buffers LBM_Domain::allocate creates, src/lbm.cpp:121-129) ---- */
typedef struct fx3d_lattice {
	int device;
	uint32_t Nx, Ny, Nz;       /* local lattice size, halo layers included (def_Nx, def_Ny, def_Nz) */
	uint32_t Dx, Dy, Dz;       /* domains per axis; >1 means the axis carries a 1-cell halo (def_Dx..) */
	uint32_t velocity_set;     /* 19 or 27 */
	uint32_t collision;        /* FX3D_SRT | FX3D_TRT */
	uint32_t storage;          /* FX3D_FP32 | FX3D_FP16S | FX3D_FP16C */
	uint32_t features;         /* FX3D_VOLUME_FORCE | FX3D_EQUILIBRIUM_BOUNDARIES | FX3D_UPDATE_FIELDS | FX3D_SUBGRID | FX3D_MOVING_BOUNDARIES */
	float w;                   /* def_w = 1/tau exactly as the device must see it (see fx3d_relaxation_rate) */
	void* fi;                  /* DDFs, device only, fx3d_fi_bytes() bytes, library-private padded SoA layout */
	float* rho;                /* [N] */
	float* u;                  /* [3N] SoA */
	uint8_t* flags;            /* [N] */
	float* F;                  /* [3N] SoA, FORCE_FIELD lattices only (LBM_Domain::F, lbm.cpp:132); else NULL */
} fx3d_lattice;

size_t fx3d_fi_bytes(const fx3d_lattice* lattice);                    /* size of the DDF buffer to allocate */
float fx3d_relaxation_rate(float nu);                                 /* def_w incl. the reference's decimal round trip (lbm.cpp:367, utilities.hpp:2745-2754) */
uint32_t fx3d_bytes_per_cell_per_step(const fx3d_lattice* lattice);   /* bandwidth_bytes_per_cell_device(), lbm.cpp:52 */

/* kernels: replace Kernel("initialize"|"stream_collide"|"update_fields") + enqueue_run (lbm.cpp:127-129,178-191) */
int fx3d_initialize(const fx3d_lattice* lattice, fx3d_stream stream);
int fx3d_stream_collide(const fx3d_lattice* lattice, uint64_t t, float fx, float fy, float fz, int region, fx3d_stream stream);
int fx3d_update_fields(const fx3d_lattice* lattice, uint64_t t, float fx, float fy, float fz, fx3d_stream stream);
/* MOVING_BOUNDARIES: re-mark the cells next to TYPE_S cells with non-zero velocity as TYPE_MS after the boundary velocities
 * changed (kernel update_moving_boundaries, kernel.cpp:1432-1450; LBM::update_moving_boundaries, lbm.cpp:1018-1027) */
int fx3d_update_moving_boundaries(const fx3d_lattice* lattice, fx3d_stream stream);
/* FORCE_FIELD (kernels update_force_field / reset_force_field / object_center_of_mass / object_force / object_torque, kernel.cpp:1873-1959;
 * LBM_Domain::enqueue_update_force_field.. lbm.cpp:206-239): update_force_field writes the force of the fluid on every TYPE_S cell into lattice->F
 * (boundary forces for lift/drag); with VOLUME_FORCE, stream_collide and update_fields add lattice->F to (fx,fy,fz) cell by cell. The object sums add
 * position, force or torque about (cx,cy,cz) over the cells whose flag byte EQUALS flag_marker into object_sum (DEVICE, 4 floats: x, y, z, cell count
 * as raw bits). Where the reference adds its work-group partial sums with floating-point atomics in arbitrary order, the order here is fixed
 * (ascending), so results are reproducible. scratch: device buffer of fx3d_object_scratch_bytes(). */
int fx3d_update_force_field(const fx3d_lattice* lattice, uint64_t t, fx3d_stream stream);
int fx3d_reset_force_field(const fx3d_lattice* lattice, fx3d_stream stream);
size_t fx3d_object_scratch_bytes(const fx3d_lattice* lattice);
int fx3d_object_center_of_mass(const fx3d_lattice* lattice, uint8_t flag_marker, float* object_sum, void* scratch, fx3d_stream stream);
int fx3d_object_force(const fx3d_lattice* lattice, uint8_t flag_marker, float* object_sum, void* scratch, fx3d_stream stream);
int fx3d_object_torque(const fx3d_lattice* lattice, uint8_t flag_marker, float cx, float cy, float cz, float* object_sum, void* scratch, fx3d_stream stream);
/* stream_collide over every non-halo cell WITH the y/z part of the halo exchange fused into it (replaces communicate_fi for those axes,
 * lbm.cpp:1355-1387): each DDF row a step writes goes straight to the memory of the domain that reads it in the next step -- this domain's,
 * or a y/z/diagonal neighbour's over NVLink. fi_neighbours[(dy+1)+3*(dz+1)] = DDF buffer of the domain at offset (dy,dz) in the domain grid
 * (periodic), identical geometry; entries for undecomposed axes and [4] (self) are ignored. The neighbours must not run step t+1 before this
 * call has finished and vice versa (one fx3d_rendezvous per step); x halos still travel with fx3d_transfer_* / fx3d_exchange_fi afterwards.
 * Only the whole-row bulk-copy kernel does this: fx3d_fused_halo_supported() tells whether a lattice qualifies (non-halo row length a multiple
 * of 4 and at most 512 cells, rows 16-byte multiples, y extent a multiple of the rows per tile). */
int fx3d_stream_collide_fused(const fx3d_lattice* lattice, uint64_t t, float fx, float fy, float fz, void* const* fi_neighbours, fx3d_stream stream);
int fx3d_fused_halo_supported(const fx3d_lattice* lattice); /* 1 or 0 */
/* n consecutive stream_collide steps t0..t0+n-1 of a single (non-decomposed) domain, no host work in between */
int fx3d_run_steps(const fx3d_lattice* lattice, uint64_t t0, uint64_t steps, float fx, float fy, float fz, fx3d_stream stream);
/* kernel choice for tests and profiling: 0 library default (persistent kernels: TMA bulk copies where the tile spans whole rows,
 * or -- FP32 -- row segments; else a cp.async ring), 1 general one-cell-per-thread kernel, 2 or 4 vector kernel with that many
 * cells per thread (falls back when the row length does not divide), 8 persistent kernel with the cp.async ring only,
 * 16 persistent kernel with bulk copies wherever they are eligible, 32 one cell per thread at high occupancy (64 registers, any grid
 * size; the reference kernel's own shape). Results are bit-identical for every choice. */
int fx3d_set_kernel_variant(int variant);
/* FX3D_REGION_INTERIOR launches of the persistent kernel leave this many resident-block slots free, so that the halo exchange
 * kernels enqueued on another stream find room beside it (default 8; 0 = occupy every slot) */
int fx3d_set_interior_reserve(int blocks);
/* stream_collide launches so far by kernel kind: 0 general (1 cell/thread), 1 vector (2/4 cells/thread), 2 persistent with a
 * cp.async ring, 3 persistent with bulk copies of whole rows, 4 persistent with bulk copies of row segments, 5 persistent with bulk loads
 * of row segments and direct stores, 6 one cell per thread at high occupancy, 7 unused (a form of kind 3 with several compute groups
 * sharing one ring of stages: measured slower and removed; kind 3 carries the fused y/z halo delivery) */
int fx3d_stream_collide_launches(int kind, uint64_t* launches);
int fx3d_launch_count(uint64_t* launches);                            /* kernels launched by this library so far */

/* GPU voxeliser (kernel voxelize_mesh / unvoxelize_mesh, kernel.cpp:2267-2357; LBM_Domain::voxelize_mesh_on_device, lbm.cpp:275-327): marks the
 * cells inside a closed triangle mesh with `flag` (and gives them the velocity of the moving/rotating body), unmarks cells the body has left.
 * p0,p1,p2: device buffers, 3 floats per triangle; bbu: HOST array of 16 floats laid out like the reference's bounding_box_and_velocity
 * (triangle count as raw bits, bounding box min-2/max+2, rotation centre, linear velocity, rotational velocity); Ox,Oy,Oz: offset of this
 * domain in the global grid (x*Nx/Dx-Hx, lbm.cpp:733); direction: axis the rays are cast along; t: the step the restarted DDFs are laid out for
 * (the reference passes t+1). */
int fx3d_voxelize_mesh(const fx3d_lattice* lattice, int Ox, int Oy, int Oz, uint32_t direction, uint64_t t, uint8_t flag,
	const float* p0, const float* p1, const float* p2, const float* bbu, fx3d_stream stream);
int fx3d_unvoxelize_mesh(const fx3d_lattice* lattice, int Ox, int Oy, int Oz, uint8_t flag, float x0, float y0, float z0, float x1, float y1, float z1, fx3d_stream stream);

/* halo transfer through linear buffers, layout and semantics of transfer_extract_fi / transfer__insert_fi and
 * transfer_*_rho_u_flags (kernel.cpp:2049-2158, lbm.cpp:1308-1354). axis: 0 x, 1 y, 2 z. Buffers are device memory of
 * fx3d_transfer_bytes() each. */
size_t fx3d_transfer_bytes(const fx3d_lattice* lattice);             /* Amax*max(transfers*sizeof(fpxx),17), lbm.cpp:1309-1315 */
int fx3d_transfer_extract_fi(const fx3d_lattice* lattice, uint32_t axis, uint64_t t, void* buf_p, void* buf_m, fx3d_stream stream);
int fx3d_transfer_insert_fi(const fx3d_lattice* lattice, uint32_t axis, uint64_t t, const void* buf_p, const void* buf_m, fx3d_stream stream);
int fx3d_transfer_extract_rho_u_flags(const fx3d_lattice* lattice, uint32_t axis, uint64_t t, void* buf_p, void* buf_m, fx3d_stream stream);
int fx3d_transfer_insert_rho_u_flags(const fx3d_lattice* lattice, uint32_t axis, uint64_t t, const void* buf_p, const void* buf_m, fx3d_stream stream);
/* the flags alone, one byte per face cell (transfer_extract_flags / transfer__insert_flags, kernel.cpp:2160-2171), and the force field, three float
 * planes per side (transfer_extract_F / transfer__insert_F, kernel.cpp:2173-2196) */
int fx3d_transfer_extract_flags(const fx3d_lattice* lattice, uint32_t axis, uint64_t t, void* buf_p, void* buf_m, fx3d_stream stream);
int fx3d_transfer_insert_flags(const fx3d_lattice* lattice, uint32_t axis, uint64_t t, const void* buf_p, const void* buf_m, fx3d_stream stream);
int fx3d_transfer_extract_F(const fx3d_lattice* lattice, uint32_t axis, uint64_t t, void* buf_p, void* buf_m, fx3d_stream stream);
int fx3d_transfer_insert_F(const fx3d_lattice* lattice, uint32_t axis, uint64_t t, const void* buf_p, const void* buf_m, fx3d_stream stream);

/* direct peer halo exchange of one axis, replacing LBM::communicate_field (lbm.cpp:1355-1383): this domain pulls what
 * the extract/swap/insert sequence would have delivered straight out of its +axis / -axis neighbours' buffers
 * (peer-mapped over NVLink, IPC-mapped from another process, or on the same GPU). The neighbours must have finished
 * the producing step; see fx3d_rendezvous_*. */
int fx3d_exchange_fi(const fx3d_lattice* lattice, uint32_t axis, uint64_t t, const void* fi_plus, const void* fi_minus, fx3d_stream stream);
int fx3d_exchange_rho_u_flags(const fx3d_lattice* lattice, uint32_t axis,
	const float* rho_plus, const float* u_plus, const uint8_t* flags_plus,
	const float* rho_minus, const float* u_minus, const uint8_t* flags_minus, fx3d_stream stream);
int fx3d_exchange_flags(const fx3d_lattice* lattice, uint32_t axis, const uint8_t* flags_plus, const uint8_t* flags_minus, fx3d_stream stream); /* communicate_flags, lbm.cpp:1391-1393 */
int fx3d_exchange_F(const fx3d_lattice* lattice, uint32_t axis, const float* F_plus, const float* F_minus, fx3d_stream stream);                /* communicate_F, lbm.cpp:1395-1397 */

/* device-side rendezvous between domains (replaces the finish_queue barriers of lbm.cpp:1357,1366,1375): each domain
 * owns an array of 64-bit counters in device memory, one per peer. signal stores `value` into slot `my_index` of every
 * listed peer array (system-scope release); wait spins on the local array until every listed slot is >= value
 * (system-scope acquire), bounded by timeout_ms (then the stream continues and the next fx3d_rendezvous_check fails). */
int fx3d_rendezvous_signal(int device, uint64_t* const* peer_arrays, int n_peers, int my_index, uint64_t value, fx3d_stream stream);
int fx3d_rendezvous_wait(int device, uint64_t* my_array, const int* peer_indices, int n_peers, uint64_t value, int timeout_ms, fx3d_stream stream);
int fx3d_rendezvous_check(int device, uint64_t* my_array, int n_slots); /* blocking; FX3D_ERR_TIMEOUT if a wait expired */

/* storage codec entry points (bit-exactness tests; device buffers) */
int fx3d_codec_encode(int device, int storage, const float* in, uint16_t* out, size_t count, fx3d_stream stream);
int fx3d_codec_decode(int device, int storage, const uint16_t* in, float* out, size_t count, fx3d_stream stream);
int fx3d_codec_fp16c_exhaustive(int device, uint64_t* mismatches, uint32_t* first_bad_bits); /* blocking, all 2^32 inputs */
/* the kernels divide by rho with a shared-reciprocal sequence; this compares it with IEEE division on `samples` operands */
int fx3d_selftest_division(int device, uint64_t samples, uint64_t* mismatches);
/* the vector kernel collides two cells per packed binary32x2 instruction; this compares every packed routine with its scalar twin */
int fx3d_selftest_packed_math(int device, uint64_t samples, uint64_t* mismatches);

#ifdef __cplusplus
}
#endif
#endif
