// lbm.hpp -- host surface of the LBM hot path: LBM_Domain (one device's share of the lattice), LBM (the global grid,
// D = Dx*Dy*Dz domains) and LBM::Memory_Container<T> (global-index view of the per-domain host buffers).
// Source-compatible with the reference's classes of the same names (FluidX3D v3.7 src/lbm.hpp:20-204, 208-611) for
// everything a setup.cpp scene on the hot path touches: constructors, run()/update_fields()/reset(), rho/u/flags with
// [] and .x/.y/.z, read_from_device()/write_to_device(), geometry helpers and getters. Device work goes through the C ABI
// of libfx3d_cuda.so; the reference's host-staged halo exchange (src/lbm.cpp:1355-1383) is replaced by direct peer
// pulls ordered by device-side rendezvous counters, and run() does not synchronise with the host between steps.
#pragma once
#include "defines.hpp"
#include "cuda.hpp"
#include "units.hpp"
#include "info.hpp"
#include "mesh.hpp"

uint bytes_per_cell_host();              // host memory per cell: rho, u, flags
uint bytes_per_cell_device();            // device memory per cell: fi, rho, u, flags
uint bandwidth_bytes_per_cell_device();  // memory traffic per cell per step: 2*Q*sizeof(fpxx)+1 (+16 with UPDATE_FIELDS)
uint3 resolution(const float3 box_aspect_ratio, const uint memory); // largest grid of the given aspect ratio that fits `memory` MB
string default_filename(const string& path, const string& name, const string& extension, const ulong t);

class LBM_Domain {
	uint Nx=1u, Ny=1u, Nz=1u; // local size, halo layers included
	uint Dx=1u, Dy=1u, Dz=1u;
	int Ox=0, Oy=0, Oz=0;     // offset of this domain in the global grid
	ulong t = 0ull;
	float nu = 1.0f/6.0f, fx = 0.0f, fy = 0.0f, fz = 0.0f;
	ulong t_last_update_fields = max_ulong;
#ifdef FORCE_FIELD
	ulong t_last_force_field = max_ulong;
	Memory<char> object_scratch; // per-work-group partial sums of the object_* reductions (device only)
#endif
	Device device;
	Memory<char> fi;          // DDFs exist on the device only, in the library's private layout
	fx3d_lattice lattice;
public:
	Memory<float> rho;        // density of every cell
	Memory<float> u;          // velocity of every cell (x, y, z planes)
	Memory<uchar> flags;      // flags of every cell
#ifdef FORCE_FIELD
	Memory<float> F;          // force on every cell (x, y, z planes)
	Memory<float> object_sum; // x, y, z, cell count (raw bits) of the last object_* reduction
#endif
	Memory<ulong> rendezvous; // 64 counters, written by the neighbouring domains (device side only)
	Memory<char> staging;     // x faces only: linear buffers [+x | -x] the neighbours pull from (x faces are strided in memory)
	ulong staging_bytes = 0ull;

	LBM_Domain(const Device_Info& device_info, fx3d_stream shared_stream, const uint Nx, const uint Ny, const uint Nz, const uint Dx, const uint Dy, const uint Dz, const int Ox, const int Oy, const int Oz, const float nu, const float fx, const float fy, const float fz);

	void enqueue_initialize();
	void enqueue_stream_collide(const int region=FX3D_REGION_ALL);
	void enqueue_stream_collide_fused(void* const* fi_neighbours); // with the y/z halo rows delivered by the kernel itself
	void enqueue_run_steps(const ulong steps); // D==1 only: `steps` stream_collide launches without host work in between
	void enqueue_update_fields();
#ifdef MOVING_BOUNDARIES
	void enqueue_update_moving_boundaries(); // mark/unmark cells next to TYPE_S cells with velocity!=0 with TYPE_MS
#endif
#ifdef FORCE_FIELD
	void enqueue_update_force_field(); // forces of the fluid on the TYPE_S cells -> F
	void enqueue_object_center_of_mass(const uchar flag_marker=TYPE_S); // sums over all cells whose flag byte equals flag_marker -> object_sum
	void enqueue_object_force(const uchar flag_marker=TYPE_S);
	void enqueue_object_torque(const float3& rotation_center, const uchar flag_marker=TYPE_S);
	void enqueue_exchange_F(const uint axis, const LBM_Domain& plus, const LBM_Domain& minus);
#endif
	void voxelize_mesh_on_device(const Mesh* mesh, const uchar flag=TYPE_S, const float3& rotation_center=float3(0.0f), const float3& linear_velocity=float3(0.0f), const float3& rotational_velocity=float3(0.0f)); // marks the cells inside the mesh
	void enqueue_unvoxelize_mesh_on_device(const Mesh* mesh, const uchar flag=TYPE_S); // clears `flag` in the bounding box of the mesh
	void enqueue_exchange_flags(const uint axis, const LBM_Domain& plus, const LBM_Domain& minus);
	void enqueue_exchange_fi(const uint axis, const LBM_Domain& plus, const LBM_Domain& minus);
	void enqueue_pack_x_faces(); // stage my outgoing x layers (transfer_extract_fi) before the rendezvous
	void enqueue_exchange_rho_u_flags(const uint axis, const LBM_Domain& plus, const LBM_Domain& minus);
	void enqueue_rendezvous_signal(const vector<LBM_Domain*>& peers, const uint my_index, const ulong value);
	void enqueue_rendezvous_wait(const vector<uint>& peer_indices, const ulong value);
	void check_rendezvous();
	void increment_time_step(const ulong steps=1ull);
	void reset_time_step();
	void finish_queue();

	const Device& get_device() const { return device; }
	const fx3d_lattice& get_lattice() const { return lattice; }
	uint get_Nx() const { return Nx; }
	uint get_Ny() const { return Ny; }
	uint get_Nz() const { return Nz; }
	ulong get_N() const { return (ulong)Nx*(ulong)Ny*(ulong)Nz; }
	uint get_Dx() const { return Dx; }
	uint get_Dy() const { return Dy; }
	uint get_Dz() const { return Dz; }
	uint get_D() const { return Dx*Dy*Dz; }
	float get_nu() const { return nu; }
	float get_tau() const { return 3.0f*nu+0.5f; }
	float get_fx() const { return fx; }
	float get_fy() const { return fy; }
	float get_fz() const { return fz; }
	ulong get_t() const { return t; }
	uint get_velocity_set() const;
	void set_fx(const float v) { fx = v; }
	void set_fy(const float v) { fy = v; }
	void set_fz(const float v) { fz = v; }
	void set_f(const float x, const float y, const float z) { fx = x; fy = y; fz = z; }
};

class LBM {
	uint Nx=1u, Ny=1u, Nz=1u; // global size (no halos)
	uint Dx=1u, Dy=1u, Dz=1u;
	bool initialized = false;
	ulong rendezvous_count = 0ull;

	void sanity_checks_constructor(const vector<Device_Info>& device_infos, const float nu, const float fx, const float fy, const float fz, const float sigma, const float alpha, const float beta, const uint particles_N, const float particles_rho);
	void sanity_checks_initialization();
	void initialize();
	void do_time_step();
	uint neighbour(const uint d, const uint axis, const int sign) const;
	uint neighbour_yz(const uint d, const int dy, const int dz) const;
	bool fused_halo = false;           // every domain's shape qualifies for fx3d_stream_collide_fused
	void rendezvous();                 // all domains meet their face neighbours on the device
	enum Field { FIELD_FI, FIELD_RHO_U_FLAGS, FIELD_FLAGS, FIELD_F };
	void communicate_field(const Field field, const uint axes=7u); // axes: bit per axis
	void communicate_fi();
	void communicate_rho_u_flags();
	void communicate_flags(); // one byte per face cell
#ifdef FORCE_FIELD
	void communicate_F();
#endif
	void construct(const uint Nx, const uint Ny, const uint Nz, const uint Dx, const uint Dy, const uint Dz, const float nu, const float fx, const float fy, const float fz, const float sigma, const float alpha, const float beta, const uint particles_N, const float particles_rho);

public:
	template<typename T> class Memory_Container { // holds no data itself, links to the domains' Memory<T>
		LBM* lbm = nullptr;
		vector<Memory<T>*> buffers;
		string name = "";
		ulong N = 0ull;
		uint d = 1u;
		uint Nx=1u, Ny=1u, Nz=1u, Dx=1u, Dy=1u, Dz=1u, D=1u, nx=1u, ny=1u, nz=1u, Hx=0u, Hy=0u, Hz=0u;
		ulong lNx=1ull, lNy=1ull, lN=1ull;
		void cache_geometry() {
			Nx = lbm->get_Nx(); Ny = lbm->get_Ny(); Nz = lbm->get_Nz(); Dx = lbm->get_Dx(); Dy = lbm->get_Dy(); Dz = lbm->get_Dz(); D = Dx*Dy*Dz;
			nx = Nx/Dx; ny = Ny/Dy; nz = Nz/Dz; Hx = Dx>1u; Hy = Dy>1u; Hz = Dz>1u;
			lNx = (ulong)(nx+2u*Hx); lNy = (ulong)(ny+2u*Hy); lN = lNx*lNy*(ulong)(nz+2u*Hz);
		}
		T& reference(const ulong i, const uint dimension) { // global index -> owning domain's host buffer
			if(D==1u) return buffers[0]->data()[i%N+(ulong)max((uint)(i/N), dimension)*N];
			const ulong g = i%N, plane = (ulong)Nx*(ulong)Ny, r = g%plane;
			const uint x = (uint)(r%(ulong)Nx), y = (uint)(r/(ulong)Nx), z = (uint)(g/plane);
			const uint domain = x/nx+(y/ny+(z/nz)*Dy)*Dx;
			const ulong local = (ulong)(x%nx+Hx)+((ulong)(y%ny+Hy)+(ulong)(z%nz+Hz)*lNy)*lNx;
			return buffers[domain]->data()[local+(ulong)max((uint)(i/N), dimension)*lN];
		}
		void write_vtk(const string& path, const bool convert_to_si_units);
	public:
		class Pointer {
			Memory_Container* memory = nullptr;
			uint dimension = 0u;
		public:
			Pointer() {}
			Pointer(Memory_Container* m, const uint dim) : memory(m), dimension(dim) {}
			T& operator[](const ulong i) { return memory->reference(i, dimension); }
			const T& operator[](const ulong i) const { return memory->reference(i, dimension); }
		};
		Pointer x, y, z;
		Memory_Container() {}
		Memory_Container(LBM* lbm_, const vector<Memory<T>*>& buffers_, const string& name_) : lbm(lbm_), buffers(buffers_), name(name_) {
			N = lbm->get_N(); d = buffers[0]->dimensions();
			cache_geometry();
			x = Pointer(this, 0u); y = Pointer(this, d>1u ? 1u : 0u); z = Pointer(this, d>2u ? 2u : 0u);
		}
		Memory_Container& operator=(Memory_Container&& m) noexcept {
			lbm = m.lbm; buffers = m.buffers; name = m.name; N = m.N; d = m.d;
			cache_geometry();
			x = Pointer(this, 0u); y = Pointer(this, d>1u ? 1u : 0u); z = Pointer(this, d>2u ? 2u : 0u);
			return *this;
		}
		void reset(const T value=(T)0) { for(Memory<T>* b : buffers) b->reset(value); }
		ulong length() const { return N; }
		uint dimensions() const { return d; }
		ulong range() const { return N*(ulong)d; }
		ulong capacity() const { return N*(ulong)d*sizeof(T); }
		T& operator[](const ulong i) { return reference(i, 0u); }
		const T& operator[](const ulong i) const { return const_cast<Memory_Container*>(this)->reference(i, 0u); }
		T operator()(const ulong i) const { return const_cast<Memory_Container*>(this)->reference(i, 0u); }
		T operator()(const ulong i, const uint dimension) const { return const_cast<Memory_Container*>(this)->reference(i, dimension); }
		void read_from_device() {
#ifndef UPDATE_FIELDS
			if(lbm->initialized) for(uint k=0u; k<D; k++) lbm->lbm_domain[k]->enqueue_update_fields(); // make rho/u current first
#endif
			for(Memory<T>* b : buffers) b->enqueue_read_from_device();
			for(Memory<T>* b : buffers) b->finish_queue();
		}
		void write_to_device() {
			for(Memory<T>* b : buffers) b->enqueue_write_to_device();
			for(Memory<T>* b : buffers) b->finish_queue();
		}
		void write_host_to_vtk(const string& path="", const bool convert_to_si_units=true) { write_vtk(default_filename(path, name, ".vtk", lbm->get_t()), convert_to_si_units); }
		void write_device_to_vtk(const string& path="", const bool convert_to_si_units=true) { read_from_device(); write_host_to_vtk(path, convert_to_si_units); }
	};

	LBM_Domain** lbm_domain = nullptr; // one per domain, d = x+(y+z*Dy)*Dx
	Memory_Container<float> rho;
	Memory_Container<float> u;
	Memory_Container<uchar> flags;
#ifdef FORCE_FIELD
	Memory_Container<float> F;
#endif

	// same constructor forms as the reference; sigma/alpha/beta/particles exist so that existing calls compile and are
	// rejected at run time when non-zero (their extensions are not part of this build)
	LBM(const uint Nx, const uint Ny, const uint Nz, const uint Dx, const uint Dy, const uint Dz, const float nu, const float fx=0.0f, const float fy=0.0f, const float fz=0.0f, const float sigma=0.0f, const float alpha=0.0f, const float beta=0.0f, const uint particles_N=0u, const float particles_rho=0.0f);
	LBM(const uint Nx, const uint Ny, const uint Nz, const float nu, const float fx=0.0f, const float fy=0.0f, const float fz=0.0f, const float sigma=0.0f, const float alpha=0.0f, const float beta=0.0f, const uint particles_N=0u, const float particles_rho=1.0f);
	LBM(const uint3 N, const uint Dx, const uint Dy, const uint Dz, const float nu, const float fx=0.0f, const float fy=0.0f, const float fz=0.0f, const float sigma=0.0f, const float alpha=0.0f, const float beta=0.0f, const uint particles_N=0u, const float particles_rho=0.0f);
	LBM(const uint3 N, const float nu, const float fx=0.0f, const float fy=0.0f, const float fz=0.0f, const float sigma=0.0f, const float alpha=0.0f, const float beta=0.0f, const uint particles_N=0u, const float particles_rho=1.0f);
	~LBM();
	LBM(const LBM&) = delete;
	LBM& operator=(const LBM&) = delete;

	void run(const ulong steps=max_ulong, const ulong total_steps=max_ulong); // first call initialises; run(0) only initialises
	void update_fields();
#ifdef MOVING_BOUNDARIES
	void update_moving_boundaries(); // mark/unmark cells next to TYPE_S cells with velocity!=0 with TYPE_MS (call after changing boundary velocities)
#endif
	void reset();
#ifdef FORCE_FIELD
	void update_force_field(); // forces of the fluid on the TYPE_S cells -> lbm.F on the device (lbm.F.read_from_device() to look at them)
	float3 object_center_of_mass(const uchar flag_marker=TYPE_S); // of all cells whose flag byte equals flag_marker
	float3 object_force(const uchar flag_marker=TYPE_S); // total force of the fluid on those cells
	float3 object_torque(const float3& rotation_center, const uchar flag_marker=TYPE_S);
#endif
	// triangle meshes: GPU voxeliser (the cells inside the closed surface receive `flag`)
	void voxelize_mesh_on_device(const Mesh* mesh, const uchar flag=TYPE_S, const float3& rotation_center=float3(0.0f), const float3& linear_velocity=float3(0.0f), const float3& rotational_velocity=float3(0.0f));
	void unvoxelize_mesh_on_device(const Mesh* mesh, const uchar flag=TYPE_S); // only needed when the bounding box changes between re-voxelisations
	void voxelize_stl(const string& path, const float3& center, const float3x3& rotation, const float size=0.0f, const uchar flag=TYPE_S); // size 0: fit into the box; >0: longest side in cells; <0: scale factor
	void voxelize_stl(const string& path, const float3x3& rotation, const float size=0.0f, const uchar flag=TYPE_S) { voxelize_stl(path, center(), rotation, size, flag); }
	void voxelize_stl(const string& path, const float3& center, const float size=0.0f, const uchar flag=TYPE_S) { voxelize_stl(path, center, float3x3(1.0f), size, flag); }
	void voxelize_stl(const string& path, const float size=0.0f, const uchar flag=TYPE_S) { voxelize_stl(path, center(), float3x3(1.0f), size, flag); }

	uint get_Nx() const { return Nx; }
	uint get_Ny() const { return Ny; }
	uint get_Nz() const { return Nz; }
	ulong get_N() const { return (ulong)Nx*(ulong)Ny*(ulong)Nz; }
	uint get_Dx() const { return Dx; }
	uint get_Dy() const { return Dy; }
	uint get_Dz() const { return Dz; }
	uint get_D() const { return Dx*Dy*Dz; }
	float get_nu() const { return lbm_domain[0]->get_nu(); }
	float get_tau() const { return 3.0f*get_nu()+0.5f; }
	float get_Re_max() const { return 0.57735027f*sqrtf((float)Nx*(float)Nx+(float)Ny*(float)Ny+(float)Nz*(float)Nz)/get_nu(); }
	float get_fx() const { return lbm_domain[0]->get_fx(); }
	float get_fy() const { return lbm_domain[0]->get_fy(); }
	float get_fz() const { return lbm_domain[0]->get_fz(); }
	ulong get_t() const { return lbm_domain[0]->get_t(); }
	uint get_velocity_set() const { return lbm_domain[0]->get_velocity_set(); }
	void set_fx(const float v) { for(uint d=0u; d<get_D(); d++) lbm_domain[d]->set_fx(v); }
	void set_fy(const float v) { for(uint d=0u; d<get_D(); d++) lbm_domain[d]->set_fy(v); }
	void set_fz(const float v) { for(uint d=0u; d<get_D(); d++) lbm_domain[d]->set_fz(v); }
	void set_f(const float x, const float y, const float z) { set_fx(x); set_fy(y); set_fz(z); }

	void coordinates(const ulong n, uint& x, uint& y, uint& z) const { // n = x+(y+z*Ny)*Nx
		const ulong plane = (ulong)Nx*(ulong)Ny, r = n%plane;
		x = (uint)(r%(ulong)Nx); y = (uint)(r/(ulong)Nx); z = (uint)(n/plane);
	}
	ulong index(const uint x, const uint y, const uint z) const { return (ulong)x+((ulong)y+(ulong)z*(ulong)Ny)*(ulong)Nx; }
	ulong index(const uint3& c) const { return index(c.x, c.y, c.z); }
	float3 position(const uint x, const uint y, const uint z) const { return float3((float)x-0.5f*(float)Nx+0.5f, (float)y-0.5f*(float)Ny+0.5f, (float)z-0.5f*(float)Nz+0.5f); }
	float3 position(const ulong n) const { uint x, y, z; coordinates(n, x, y, z); return position(x, y, z); }
	float3 size() const { return float3((float)Nx, (float)Ny, (float)Nz); }
	float3 center() const { return float3(0.5f*(float)Nx-0.5f, 0.5f*(float)Ny-0.5f, 0.5f*(float)Nz-0.5f); }
	uint smallest_side_length() const { return min(min(Nx, Ny), Nz); }
	uint largest_side_length() const { return max(max(Nx, Ny), Nz); }
	float3 relative_position(const uint x, const uint y, const uint z) const { return float3(((float)x+0.5f)/(float)Nx-0.5f, ((float)y+0.5f)/(float)Ny-0.5f, ((float)z+0.5f)/(float)Nz-0.5f); }
	float3 relative_position(const ulong n) const { uint x, y, z; coordinates(n, x, y, z); return relative_position(x, y, z); }
	void write_status(const string& path=""); // text report of the simulation parameters
};

// ---- VTK export of a field: binary STRUCTURED_POINTS, big-endian payload, as the reference writes it ----
template<typename T> void LBM::Memory_Container<T>::write_vtk(const string& path, const bool convert_to_si_units) {
	float spacing = 1.0f;
	T factor = (T)1;
	if(convert_to_si_units) {
		spacing = units.si_x(1.0f);
		if(name=="rho") factor = (T)units.si_rho(1.0f);
		if(name=="u") factor = (T)units.si_u(1.0f);
	}
	string type = "float";
	if(std::is_same<T, uchar>::value) type = "unsigned_char";
	std::FILE* file = std::fopen(path.c_str(), "wb");
	if(!file) { print_warning("File \""+path+"\" could not be written."); return; }
	const float3 origin = spacing*float3(0.5f-0.5f*(float)Nx, 0.5f-0.5f*(float)Ny, 0.5f-0.5f*(float)Nz);
	const string header = "# vtk DataFile Version 3.0\nfx3d-b200 "+name+"\nBINARY\nDATASET STRUCTURED_POINTS\nDIMENSIONS "+to_string(Nx)+" "+to_string(Ny)+" "+to_string(Nz)+
		"\nORIGIN "+to_string(origin.x)+" "+to_string(origin.y)+" "+to_string(origin.z)+"\nSPACING "+to_string(spacing)+" "+to_string(spacing)+" "+to_string(spacing)+
		"\nPOINT_DATA "+to_string(N)+"\nSCALARS data "+type+" "+to_string(d)+"\nLOOKUP_TABLE default\n";
	std::fwrite(header.data(), 1, header.size(), file);
	vector<T> payload((size_t)range());
	for(ulong n=0ull; n<N; n++) for(uint c=0u; c<d; c++) { // interleave components, swap to big endian
		T v = reference(n, c)*factor;
		uchar bytes[sizeof(T)];
		std::memcpy(bytes, &v, sizeof(T));
		std::reverse(bytes, bytes+sizeof(T));
		std::memcpy(&payload[(size_t)(n*(ulong)d+c)], bytes, sizeof(T));
	}
	std::fwrite(payload.data(), sizeof(T), payload.size(), file);
	std::fclose(file);
	print_info("File \""+path+"\" saved.");
}
