// lbm.cpp -- implementation of LBM_Domain and LBM on top of libfx3d_cuda.so (include/fx3d.h).
// Follows the sequencing of the reference's src/lbm.cpp (FluidX3D v3.7): LBM_Domain ctor/allocate :96-176, launches
// :178-191, LBM ctor :721-773, sanity checks :783-879, initialize :881-922, do_time_step :924-953, run :955-975,
// halo exchange :1355-1390 -- with the device layer replaced and no host synchronisation between time steps.
#include "lbm.hpp"

Units units;

#if defined(D3Q19)
static const uint velocity_set = 19u, transfers = 5u;
#else
static const uint velocity_set = 27u, transfers = 9u;
#endif
#if defined(TRT)
static const uint collision_operator = FX3D_TRT;
#else
static const uint collision_operator = FX3D_SRT;
#endif
#if defined(FP16S)
static const uint storage_format = FX3D_FP16S;
#elif defined(FP16C)
static const uint storage_format = FX3D_FP16C;
#else
static const uint storage_format = FX3D_FP32;
#endif
static uint extension_mask() {
	uint m = 0u;
#ifdef VOLUME_FORCE
	m |= FX3D_VOLUME_FORCE;
#endif
#ifdef EQUILIBRIUM_BOUNDARIES
	m |= FX3D_EQUILIBRIUM_BOUNDARIES;
#endif
#ifdef UPDATE_FIELDS
	m |= FX3D_UPDATE_FIELDS;
#endif
#ifdef SUBGRID
	m |= FX3D_SUBGRID;
#endif
#ifdef MOVING_BOUNDARIES
	m |= FX3D_MOVING_BOUNDARIES;
#endif
#ifdef FORCE_FIELD
	m |= FX3D_FORCE_FIELD;
#endif
	return m;
}

#ifdef FORCE_FIELD
static const uint force_field_bytes = 12u; // F
#else
static const uint force_field_bytes = 0u;
#endif
uint bytes_per_cell_host() { return 17u+force_field_bytes; } // rho 4 + u 12 + flags 1 (+ F 12)
uint bytes_per_cell_device() { return velocity_set*(uint)sizeof(fpxx)+17u+force_field_bytes; }
uint bandwidth_bytes_per_cell_device() {
	uint b = velocity_set*2u*(uint)sizeof(fpxx)+1u+force_field_bytes; // every DDF read once and written once, plus the flag byte
#ifdef UPDATE_FIELDS
	b += 16u;
#endif
#ifdef MOVING_BOUNDARIES
	b += velocity_set-1u; // the reference counts the neighbour flags
#endif
	return b;
}
uint3 resolution(const float3 box_aspect_ratio, const uint memory) {
	const float per_unit_box = box_aspect_ratio.x*box_aspect_ratio.y*box_aspect_ratio.z*(float)bytes_per_cell_device()/1048576.0f; // MB
	const float scaling = cbrtf((float)memory/per_unit_box);
	return uint3(to_uint(scaling*box_aspect_ratio.x), to_uint(scaling*box_aspect_ratio.y), to_uint(scaling*box_aspect_ratio.z));
}
string default_filename(const string& path, const string& name, const string& extension, const ulong t) {
	char stamp[32];
	std::snprintf(stamp, sizeof(stamp), "%09llu", (unsigned long long)t);
	return (path=="" ? get_exe_path()+"export/" : path)+(name=="" ? "file" : name)+"-"+stamp+extension;
}

// ---------------------------------------------------------------- LBM_Domain ----------------------------------------------------------------

LBM_Domain::LBM_Domain(const Device_Info& device_info, fx3d_stream shared_stream, const uint Nx, const uint Ny, const uint Nz, const uint Dx, const uint Dy, const uint Dz, const int Ox, const int Oy, const int Oz, const float nu, const float fx, const float fy, const float fz)
	: Nx(Nx), Ny(Ny), Nz(Nz), Dx(Dx), Dy(Dy), Dz(Dz), Ox(Ox), Oy(Oy), Oz(Oz), nu(nu), fx(fx), fy(fy), fz(fz), device(device_info, shared_stream) {
	lattice.device = device.ordinal();
	lattice.Nx = Nx; lattice.Ny = Ny; lattice.Nz = Nz; lattice.Dx = Dx; lattice.Dy = Dy; lattice.Dz = Dz;
	lattice.velocity_set = velocity_set; lattice.collision = collision_operator; lattice.storage = storage_format; lattice.features = extension_mask();
	lattice.w = fx3d_relaxation_rate(nu); // 1/tau with the decimal round trip the reference's JIT constant goes through
	lattice.fi = nullptr; lattice.rho = nullptr; lattice.u = nullptr; lattice.flags = nullptr; lattice.F = nullptr;
	const size_t fi_bytes = fx3d_fi_bytes(&lattice);
	if(fi_bytes==0u) print_error(string("lattice rejected: ")+fx3d_last_error());
	print_info("Allocating memory. This may take a few seconds.");
	const ulong N = get_N();
	fi = Memory<char>(device, (ulong)fi_bytes, 1u, false);   // device only
	rho = Memory<float>(device, N, 1u, true, true, 1.0f);
	u = Memory<float>(device, N, 3u);
	flags = Memory<uchar>(device, N);
	if(get_D()>1u) rendezvous = Memory<ulong>(device, 64ull, 1u, false);
	lattice.fi = fi.device_data(); lattice.rho = rho.device_data(); lattice.u = u.device_data(); lattice.flags = flags.device_data();
	if(Dx>1u) {
		staging_bytes = ((ulong)fx3d_transfer_bytes(&lattice)+255ull)/256ull*256ull;
		staging = Memory<char>(device, 2ull*staging_bytes, 1u, false);
	}
#ifdef FORCE_FIELD
	F = Memory<float>(device, N, 3u);
	object_sum = Memory<float>(device, 1ull, 4u);
	object_scratch = Memory<char>(device, (ulong)fx3d_object_scratch_bytes(&lattice), 1u, false);
	lattice.F = F.device_data();
#endif
}
uint LBM_Domain::get_velocity_set() const { return velocity_set; }

void LBM_Domain::enqueue_initialize() { fx3d_check(fx3d_initialize(&lattice, device.get_stream()), "initialize"); }
void LBM_Domain::enqueue_stream_collide_fused(void* const* fi_neighbours) { fx3d_check(fx3d_stream_collide_fused(&lattice, t, fx, fy, fz, fi_neighbours, device.get_stream()), "stream_collide_fused"); }
void LBM_Domain::enqueue_stream_collide(const int region) { fx3d_check(fx3d_stream_collide(&lattice, t, fx, fy, fz, region, device.get_stream()), "stream_collide"); }
void LBM_Domain::enqueue_run_steps(const ulong steps) { fx3d_check(fx3d_run_steps(&lattice, t, steps, fx, fy, fz, device.get_stream()), "stream_collide"); }
#ifdef MOVING_BOUNDARIES
void LBM_Domain::enqueue_update_moving_boundaries() { fx3d_check(fx3d_update_moving_boundaries(&lattice, device.get_stream()), "update_moving_boundaries"); }
#endif
#ifdef FORCE_FIELD
void LBM_Domain::enqueue_update_force_field() {
	if(t!=t_last_force_field) { // F is stale only if time has advanced since the last update
		fx3d_check(fx3d_update_force_field(&lattice, t, device.get_stream()), "update_force_field");
		t_last_force_field = t;
	}
}
void LBM_Domain::enqueue_object_center_of_mass(const uchar flag_marker) {
	fx3d_check(fx3d_object_center_of_mass(&lattice, flag_marker, object_sum.device_data(), object_scratch.device_data(), device.get_stream()), "object_center_of_mass");
	object_sum.enqueue_read_from_device();
}
void LBM_Domain::enqueue_object_force(const uchar flag_marker) {
	enqueue_update_force_field();
	fx3d_check(fx3d_object_force(&lattice, flag_marker, object_sum.device_data(), object_scratch.device_data(), device.get_stream()), "object_force");
	object_sum.enqueue_read_from_device();
}
void LBM_Domain::enqueue_object_torque(const float3& rotation_center, const uchar flag_marker) {
	enqueue_update_force_field();
	fx3d_check(fx3d_object_torque(&lattice, flag_marker, rotation_center.x, rotation_center.y, rotation_center.z, object_sum.device_data(), object_scratch.device_data(), device.get_stream()), "object_torque");
	object_sum.enqueue_read_from_device();
}
void LBM_Domain::enqueue_exchange_F(const uint axis, const LBM_Domain& plus, const LBM_Domain& minus) {
	fx3d_check(fx3d_exchange_F(&lattice, axis, plus.lattice.F, minus.lattice.F, device.get_stream()), "halo exchange (F)");
}
#endif
void LBM_Domain::voxelize_mesh_on_device(const Mesh* mesh, const uchar flag, const float3& rotation_center, const float3& linear_velocity, const float3& rotational_velocity) {
	Memory<float> p0(device, (ulong)mesh->triangle_number, 3u), p1(device, (ulong)mesh->triangle_number, 3u), p2(device, (ulong)mesh->triangle_number, 3u); // xyz interleaved per triangle
	for(uint k=0u; k<mesh->triangle_number; k++) {
		p0[3ull*k] = mesh->p0[k].x; p0[3ull*k+1ull] = mesh->p0[k].y; p0[3ull*k+2ull] = mesh->p0[k].z;
		p1[3ull*k] = mesh->p1[k].x; p1[3ull*k+1ull] = mesh->p1[k].y; p1[3ull*k+2ull] = mesh->p1[k].z;
		p2[3ull*k] = mesh->p2[k].x; p2[3ull*k+1ull] = mesh->p2[k].y; p2[3ull*k+2ull] = mesh->p2[k].z;
	}
	// bounding box with 2 cells of tolerance (re-voxelisation of moving objects), rotation centre, velocities: the kernel's parameter block
	const float x0 = mesh->pmin.x-2.0f, y0 = mesh->pmin.y-2.0f, z0 = mesh->pmin.z-2.0f, x1 = mesh->pmax.x+2.0f, y1 = mesh->pmax.y+2.0f, z1 = mesh->pmax.z+2.0f;
	const float block[16] = { as_float(mesh->triangle_number), x0, y0, z0, x1, y1, z1, rotation_center.x, rotation_center.y, rotation_center.z,
		linear_velocity.x, linear_velocity.y, linear_velocity.z, rotational_velocity.x, rotational_velocity.y, rotational_velocity.z };
	uint direction = 0u; // rays along the rotation axis, or -- for a body that does not rotate -- through the smallest face of the bounding box
	if(length(rotational_velocity)==0.0f) {
		const float area[3] = { (y1-y0)*(z1-z0), (z1-z0)*(x1-x0), (x1-x0)*(y1-y0) };
		for(uint i=1u; i<3u; i++) if(area[i]<area[direction]) direction = i;
	} else {
		const float along[3] = { fabsf(rotational_velocity.x), fabsf(rotational_velocity.y), fabsf(rotational_velocity.z) };
		for(uint i=1u; i<3u; i++) if(along[i]>along[direction]) direction = i;
	}
	p0.write_to_device(); p1.write_to_device(); p2.write_to_device();
	fx3d_check(fx3d_voxelize_mesh(&lattice, Ox, Oy, Oz, direction, t+1ull, flag, p0.device_data(), p1.device_data(), p2.device_data(), block, device.get_stream()), "voxelize_mesh");
	finish_queue();
}
void LBM_Domain::enqueue_unvoxelize_mesh_on_device(const Mesh* mesh, const uchar flag) {
	fx3d_check(fx3d_unvoxelize_mesh(&lattice, Ox, Oy, Oz, flag, mesh->pmin.x, mesh->pmin.y, mesh->pmin.z, mesh->pmax.x, mesh->pmax.y, mesh->pmax.z, device.get_stream()), "unvoxelize_mesh");
}
void LBM_Domain::enqueue_exchange_flags(const uint axis, const LBM_Domain& plus, const LBM_Domain& minus) {
	fx3d_check(fx3d_exchange_flags(&lattice, axis, plus.lattice.flags, minus.lattice.flags, device.get_stream()), "halo exchange (flags)");
}
void LBM_Domain::enqueue_update_fields() {
#ifndef UPDATE_FIELDS
	if(t!=t_last_update_fields) { // rho/u on the device are stale only if time has advanced since the last update
		fx3d_check(fx3d_update_fields(&lattice, t, fx, fy, fz, device.get_stream()), "update_fields");
		t_last_update_fields = t;
	}
#endif
}
void LBM_Domain::enqueue_pack_x_faces() {
	fx3d_check(fx3d_transfer_extract_fi(&lattice, 0u, t, staging.device_data(), staging.device_data()+staging_bytes, device.get_stream()), "halo exchange (pack x faces)");
}
void LBM_Domain::enqueue_exchange_fi(const uint axis, const LBM_Domain& plus, const LBM_Domain& minus) {
	if(axis==0u) // my +x halo receives what the +x neighbour packed for its -x side, and vice versa (coalesced peer reads)
		fx3d_check(fx3d_transfer_insert_fi(&lattice, 0u, t, plus.staging.device_data()+plus.staging_bytes, minus.staging.device_data(), device.get_stream()), "halo exchange (fi, x)");
	else
		fx3d_check(fx3d_exchange_fi(&lattice, axis, t, plus.lattice.fi, minus.lattice.fi, device.get_stream()), "halo exchange (fi)");
}
void LBM_Domain::enqueue_exchange_rho_u_flags(const uint axis, const LBM_Domain& plus, const LBM_Domain& minus) {
	fx3d_check(fx3d_exchange_rho_u_flags(&lattice, axis, plus.lattice.rho, plus.lattice.u, plus.lattice.flags, minus.lattice.rho, minus.lattice.u, minus.lattice.flags, device.get_stream()), "halo exchange (rho, u, flags)");
}
void LBM_Domain::enqueue_rendezvous_signal(const vector<LBM_Domain*>& peers, const uint my_index, const ulong value) {
	vector<uint64_t*> arrays;
	for(LBM_Domain* p : peers) arrays.push_back((uint64_t*)p->rendezvous.device_data());
	fx3d_check(fx3d_rendezvous_signal(device.ordinal(), arrays.data(), (int)arrays.size(), (int)my_index, value, device.get_stream()), "rendezvous signal");
}
void LBM_Domain::enqueue_rendezvous_wait(const vector<uint>& peer_indices, const ulong value) {
	vector<int> idx(peer_indices.begin(), peer_indices.end());
	fx3d_check(fx3d_rendezvous_wait(device.ordinal(), (uint64_t*)rendezvous.device_data(), idx.data(), (int)idx.size(), value, 20000, device.get_stream()), "rendezvous wait");
}
void LBM_Domain::check_rendezvous() { fx3d_check(fx3d_rendezvous_check(device.ordinal(), (uint64_t*)rendezvous.device_data(), 64), "halo exchange"); }
void LBM_Domain::increment_time_step(const ulong steps) {
	t += steps;
#ifdef UPDATE_FIELDS
	t_last_update_fields = t;
#endif
}
void LBM_Domain::reset_time_step() {
	t = 0ull;
#ifdef UPDATE_FIELDS
	t_last_update_fields = t;
#endif
}
void LBM_Domain::finish_queue() { device.finish_queue(); }

// -------------------------------------------------------------------- LBM --------------------------------------------------------------------

extern vector<string> main_arguments; // device IDs given on the command line (main.cpp)

static vector<Device_Info> smart_device_selection(const uint D) { // D devices for D domains; else one device hosts all domains
	const vector<Device_Info> devices = get_devices();
	vector<Device_Info> chosen(D);
	if((uint)main_arguments.size()==D) {
		for(uint d=0u; d<D; d++) chosen[d] = select_device_with_id((uint)std::stoul(main_arguments[d]), devices);
	} else if(main_arguments.size()>0u) {
		print_warning("Incorrect number of devices specified. Using single fastest device for all domains.");
		for(uint d=0u; d<D; d++) chosen[d] = select_device_with_most_flops(devices);
	} else if((uint)devices.size()>=D) {
		for(uint d=0u; d<D; d++) chosen[d] = devices[d]; // all CUDA devices of one node are of one type in practice
	} else {
		print_warning("Not enough devices of the same type available. Using single fastest device for all domains.");
		for(uint d=0u; d<D; d++) chosen[d] = select_device_with_most_flops(devices);
	}
	return chosen;
}

void LBM::construct(const uint Nx_, const uint Ny_, const uint Nz_, const uint Dx_, const uint Dy_, const uint Dz_, const float nu, const float fx, const float fy, const float fz, const float sigma, const float alpha, const float beta, const uint particles_N, const float particles_rho) {
	if(Dx_*Dy_*Dz_==0u) print_error("You specified 0 LBM grid domains ("+to_string(Dx_)+"x"+to_string(Dy_)+"x"+to_string(Dz_)+"). There has to be at least 1 domain in every direction. Check your input in LBM constructor.");
	const uint NDx = (Nx_/Dx_)*Dx_, NDy = (Ny_/Dy_)*Dy_, NDz = (Nz_/Dz_)*Dz_; // equal domains only
	if(NDx!=Nx_||NDy!=Ny_||NDz!=Nz_) print_warning("LBM grid ("+to_string(Nx_)+"x"+to_string(Ny_)+"x"+to_string(Nz_)+") is not equally divisible in domains ("+to_string(Dx_)+"x"+to_string(Dy_)+"x"+to_string(Dz_)+"). Changing resolution to ("+to_string(NDx)+"x"+to_string(NDy)+"x"+to_string(NDz)+").");
	Nx = NDx; Ny = NDy; Nz = NDz; Dx = Dx_; Dy = Dy_; Dz = Dz_;
	const uint D = get_D(), Hx = Dx>1u, Hy = Dy>1u, Hz = Dz>1u;
	if(D>63u) print_error("At most 63 domains are supported.");
	const vector<Device_Info> device_infos = smart_device_selection(D);
	sanity_checks_constructor(device_infos, nu, fx, fy, fz, sigma, alpha, beta, particles_N, particles_rho);
	lbm_domain = new LBM_Domain*[D];
	vector<fx3d_stream> stream_of_device(64, nullptr); // domains that share a physical device share its in-order stream
	for(uint d=0u; d<D; d++) {
		const uint x = (d%(Dx*Dy))%Dx, y = (d%(Dx*Dy))/Dx, z = d/(Dx*Dy);
		const uint id = device_infos[d].id;
		lbm_domain[d] = new LBM_Domain(device_infos[d], id<64u ? stream_of_device[id] : nullptr, Nx/Dx+2u*Hx, Ny/Dy+2u*Hy, Nz/Dz+2u*Hz, Dx, Dy, Dz,
			(int)(x*Nx/Dx)-(int)Hx, (int)(y*Ny/Dy)-(int)Hy, (int)(z*Nz/Dz)-(int)Hz, nu, fx, fy, fz);
		if(id<64u && !stream_of_device[id]) stream_of_device[id] = lbm_domain[d]->get_device().get_stream();
	}
	for(uint a=0u; a<D; a++) for(uint b=0u; b<D; b++) { // neighbours read each other's buffers directly
		const int da = lbm_domain[a]->get_device().ordinal(), db = lbm_domain[b]->get_device().ordinal();
		if(da!=db) fx3d_check(fx3d_device_enable_peer(da, db), "peer access");
	}
	vector<Memory<float>*> b_rho, b_u; vector<Memory<uchar>*> b_flags;
	for(uint d=0u; d<D; d++) { b_rho.push_back(&lbm_domain[d]->rho); b_u.push_back(&lbm_domain[d]->u); b_flags.push_back(&lbm_domain[d]->flags); }
	rho = Memory_Container<float>(this, b_rho, "rho");
	u = Memory_Container<float>(this, b_u, "u");
	flags = Memory_Container<uchar>(this, b_flags, "flags");
#ifdef FORCE_FIELD
	vector<Memory<float>*> b_F;
	for(uint d=0u; d<D; d++) b_F.push_back(&lbm_domain[d]->F);
	F = Memory_Container<float>(this, b_F, "F");
#endif
	fused_halo = Dy*Dz>1u;
	for(uint d=0u; d<D && fused_halo; d++) fused_halo = fx3d_fused_halo_supported(&lbm_domain[d]->get_lattice())==1;
}
LBM::LBM(const uint Nx, const uint Ny, const uint Nz, const uint Dx, const uint Dy, const uint Dz, const float nu, const float fx, const float fy, const float fz, const float sigma, const float alpha, const float beta, const uint particles_N, const float particles_rho) {
	construct(Nx, Ny, Nz, Dx, Dy, Dz, nu, fx, fy, fz, sigma, alpha, beta, particles_N, particles_rho);
}
LBM::LBM(const uint Nx, const uint Ny, const uint Nz, const float nu, const float fx, const float fy, const float fz, const float sigma, const float alpha, const float beta, const uint particles_N, const float particles_rho) {
	construct(Nx, Ny, Nz, 1u, 1u, 1u, nu, fx, fy, fz, sigma, alpha, beta, particles_N, particles_rho==1.0f ? 0.0f : particles_rho);
}
LBM::LBM(const uint3 N, const uint Dx, const uint Dy, const uint Dz, const float nu, const float fx, const float fy, const float fz, const float sigma, const float alpha, const float beta, const uint particles_N, const float particles_rho) {
	construct(N.x, N.y, N.z, Dx, Dy, Dz, nu, fx, fy, fz, sigma, alpha, beta, particles_N, particles_rho);
}
LBM::LBM(const uint3 N, const float nu, const float fx, const float fy, const float fz, const float sigma, const float alpha, const float beta, const uint particles_N, const float particles_rho) {
	construct(N.x, N.y, N.z, 1u, 1u, 1u, nu, fx, fy, fz, sigma, alpha, beta, particles_N, particles_rho==1.0f ? 0.0f : particles_rho);
}
LBM::~LBM() {
	info.print_finalize();
	for(uint d=0u; d<get_D(); d++) delete lbm_domain[d];
	delete[] lbm_domain;
}

void LBM::sanity_checks_constructor(const vector<Device_Info>& device_infos, const float nu, const float fx, const float fy, const float fz, const float sigma, const float alpha, const float beta, const uint particles_N, const float particles_rho) {
	if((ulong)Nx*(ulong)Ny*(ulong)Nz==0ull) print_error("Grid point number is 0: "+to_string(Nx)+"x"+to_string(Ny)+"x"+to_string(Nz)+" = 0.");
	uint memory_available = max_uint;
	for(const Device_Info& i : device_infos) memory_available = min(memory_available, i.memory);
	const uint memory_required = (uint)(get_N()/(ulong)get_D()*(ulong)bytes_per_cell_device()/1048576ull);
	if(memory_required>memory_available) {
		const float factor = cbrtf((float)memory_available/(float)memory_required);
		string message = "Grid resolution ("+to_string(Nx)+", "+to_string(Ny)+", "+to_string(Nz)+") is too large: "+to_string(get_D())+"x "+to_string(memory_required)+" MB required, "+to_string(get_D())+"x "+to_string(memory_available)+" MB available. Largest possible resolution is ("+to_string((uint)(factor*(float)Nx))+", "+to_string((uint)(factor*(float)Ny))+", "+to_string((uint)(factor*(float)Nz))+").";
#if !defined(FP16S)&&!defined(FP16C)
		message += " Consider using FP16S/FP16C memory compression to double the maximum grid resolution; for this, uncomment \"#define FP16S\" or \"#define FP16C\" in defines.hpp.";
#endif
		print_error(message);
	}
	if(nu==0.0f) print_error("Viscosity cannot be 0. Change it in setup.cpp.");
	else if(nu<0.0f) print_error("Viscosity cannot be negative. Remove the \"-\" in setup.cpp.");
#if !defined(SRT)&&!defined(TRT)
	print_error("No LBM collision operator selected. Uncomment either \"#define SRT\" or \"#define TRT\" in defines.hpp");
#elif defined(SRT)&&defined(TRT)
	print_error("Too many LBM collision operators selected. Comment out either \"#define SRT\" or \"#define TRT\" in defines.hpp");
#endif
#ifndef VOLUME_FORCE
	if(fx!=0.0f||fy!=0.0f||fz!=0.0f) print_error("Volume force is set in LBM constructor in main_setup(), but VOLUME_FORCE is not enabled. Uncomment \"#define VOLUME_FORCE\" in defines.hpp.");
#else
	if(fx==0.0f&&fy==0.0f&&fz==0.0f) print_warning("The VOLUME_FORCE extension is enabled but the volume force in LBM constructor is set to zero. You may disable the extension by commenting out \"#define VOLUME_FORCE\" in defines.hpp.");
#endif
	if(sigma!=0.0f) print_error("Surface tension is set in LBM constructor in main_setup(), but SURFACE is not part of this build.");
	if(alpha!=0.0f||beta!=0.0f) print_error("Thermal diffusion/expansion coefficients are set in LBM constructor in main_setup(), but TEMPERATURE is not part of this build.");
	if(particles_N>0u) print_error("The number of particles is set to "+to_string(particles_N)+">0, but PARTICLES is not part of this build.");
	(void)particles_rho;
}
void LBM::sanity_checks_initialization() { // which extensions do the flags call for?
	const uint threads = max(1u, (uint)thread::hardware_concurrency());
	vector<char> uses_e(threads, 0), moving(threads, 0);
	vector<uchar> any(threads, (uchar)0);
	parallel_for(get_N(), threads, [&](ulong n, uint t) {
		const uchar f = flags[n], bo = f&(TYPE_S|TYPE_E);
		any[t] = any[t]|f;
		if(bo==TYPE_E) uses_e[t] = 1;
		if(bo&TYPE_S) if((bo==TYPE_S&&(u.x[n]!=0.0f||u.y[n]!=0.0f||u.z[n]!=0.0f))||bo==(TYPE_S|TYPE_E)) moving[t] = 1;
	});
	bool e = false, m = false; uchar used = 0u;
	for(uint t=0u; t<threads; t++) { e = e||uses_e[t]; m = m||moving[t]; used = used|any[t]; }
#ifndef MOVING_BOUNDARIES
	if(m) print_warning("Some boundary cells have non-zero velocity, but MOVING_BOUNDARIES is not enabled.");
#endif
#ifndef EQUILIBRIUM_BOUNDARIES
	if(e) print_error("Some cells are set as equilibrium boundaries with the TYPE_E flag, but EQUILIBRIUM_BOUNDARIES is not enabled. Uncomment \"#define EQUILIBRIUM_BOUNDARIES\" in defines.hpp.");
#else
	if(!e) print_warning("The EQUILIBRIUM_BOUNDARIES extension is enabled but no equilibrium boundary cells (TYPE_E flag) are placed in the simulation box. You may disable the extension by commenting out \"#define EQUILIBRIUM_BOUNDARIES\" in defines.hpp.");
#endif
	if(used&(TYPE_F|TYPE_I|TYPE_G)) print_error("Some cells are set as fluid/interface/gas with the TYPE_F/TYPE_I/TYPE_G flags, but SURFACE is not part of this build.");
	if(used&TYPE_T) print_error("Some cells are set as temperature boundary with the TYPE_T flag, but TEMPERATURE is not part of this build.");
}

uint LBM::neighbour(const uint d, const uint axis, const int sign) const {
	uint c[3] = { (d%(Dx*Dy))%Dx, (d%(Dx*Dy))/Dx, d/(Dx*Dy) };
	const uint Dn[3] = { Dx, Dy, Dz };
	c[axis] = (c[axis]+Dn[axis]+(uint)sign)%Dn[axis];
	return c[0]+(c[1]+c[2]*Dy)*Dx;
}
uint LBM::neighbour_yz(const uint d, const int dy, const int dz) const {
	const uint x = (d%(Dx*Dy))%Dx, y = (d%(Dx*Dy))/Dx, z = d/(Dx*Dy);
	return x+((y+Dy+(uint)dy)%Dy+((z+Dz+(uint)dz)%Dz)*Dy)*Dx;
}
void LBM::rendezvous() { // every domain tells its face neighbours "I am here" and waits for theirs, on the device
	rendezvous_count++;
	const uint Dn[3] = { Dx, Dy, Dz };
	vector<vector<uint>> nb(get_D());
	for(uint d=0u; d<get_D(); d++) {
		for(uint axis=0u; axis<3u; axis++) if(Dn[axis]>1u) for(int sign=-1; sign<=1; sign+=2) {
			const uint n = neighbour(d, axis, sign);
			if(n!=d && std::find(nb[d].begin(), nb[d].end(), n)==nb[d].end()) nb[d].push_back(n);
		}
		if(Dy>1u && Dz>1u) for(int dy=-1; dy<=1; dy+=2) for(int dz=-1; dz<=1; dz+=2) { // with fused y/z halo delivery the diagonal neighbours write into this domain too
			const uint n = neighbour_yz(d, dy, dz);
			if(n!=d && std::find(nb[d].begin(), nb[d].end(), n)==nb[d].end()) nb[d].push_back(n);
		}
		std::sort(nb[d].begin(), nb[d].end());
		vector<LBM_Domain*> peers;
		for(uint n : nb[d]) peers.push_back(lbm_domain[n]);
		lbm_domain[d]->enqueue_rendezvous_signal(peers, d, rendezvous_count);
	}
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->enqueue_rendezvous_wait(nb[d], rendezvous_count);
}
void LBM::communicate_field(const Field field, const uint axes) { // x, then y, then z, so that edges and corners travel with later faces
	const uint Dn[3] = { Dx, Dy, Dz };
	for(uint axis=0u; axis<3u; axis++) if(Dn[axis]>1u && ((axes>>axis)&1u)) {
		if(field==FIELD_FI && axis==0u) for(uint d=0u; d<get_D(); d++) lbm_domain[d]->enqueue_pack_x_faces();
		rendezvous(); // what this phase reads (stream_collide output, or the previous axis' halos) is complete everywhere
		for(uint d=0u; d<get_D(); d++) {
			LBM_Domain& plus = *lbm_domain[neighbour(d, axis, +1)]; LBM_Domain& minus = *lbm_domain[neighbour(d, axis, -1)];
			if(field==FIELD_FI) lbm_domain[d]->enqueue_exchange_fi(axis, plus, minus);
			else if(field==FIELD_FLAGS) lbm_domain[d]->enqueue_exchange_flags(axis, plus, minus);
#ifdef FORCE_FIELD
			else if(field==FIELD_F) lbm_domain[d]->enqueue_exchange_F(axis, plus, minus);
#endif
			else lbm_domain[d]->enqueue_exchange_rho_u_flags(axis, plus, minus);
		}
	}
	rendezvous(); // nobody overwrites what a neighbour is still reading
}
#ifdef MOVING_BOUNDARIES
void LBM::update_moving_boundaries() { // src/lbm.cpp:1018-1027
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->enqueue_update_moving_boundaries();
	if(get_D()>1u) communicate_flags();
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->finish_queue();
}
#endif
void LBM::communicate_fi() { communicate_field(FIELD_FI); }
void LBM::communicate_rho_u_flags() { communicate_field(FIELD_RHO_U_FLAGS); }
void LBM::communicate_flags() { communicate_field(FIELD_FLAGS); }
#ifdef FORCE_FIELD
void LBM::communicate_F() { communicate_field(FIELD_F); }
void LBM::update_force_field() {
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->enqueue_update_force_field();
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->finish_queue();
}
float3 LBM::object_center_of_mass(const uchar flag_marker) { // per-domain sums, added in domain order
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->enqueue_object_center_of_mass(flag_marker);
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->finish_queue();
	float3 sum(0.0f); ulong cells = 0ull;
	for(uint d=0u; d<get_D(); d++) { sum += float3(lbm_domain[d]->object_sum.x[0], lbm_domain[d]->object_sum.y[0], lbm_domain[d]->object_sum.z[0]); cells += (ulong)as_uint(lbm_domain[d]->object_sum.w[0]); }
	return sum/(float)cells;
}
float3 LBM::object_force(const uchar flag_marker) {
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->enqueue_object_force(flag_marker);
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->finish_queue();
	float3 sum(0.0f);
	for(uint d=0u; d<get_D(); d++) sum += float3(lbm_domain[d]->object_sum.x[0], lbm_domain[d]->object_sum.y[0], lbm_domain[d]->object_sum.z[0]);
	return sum;
}
float3 LBM::object_torque(const float3& rotation_center, const uchar flag_marker) {
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->enqueue_object_torque(rotation_center, flag_marker);
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->finish_queue();
	float3 sum(0.0f);
	for(uint d=0u; d<get_D(); d++) sum += float3(lbm_domain[d]->object_sum.x[0], lbm_domain[d]->object_sum.y[0], lbm_domain[d]->object_sum.z[0]);
	return sum;
}
#endif
void LBM::voxelize_mesh_on_device(const Mesh* mesh, const uchar flag, const float3& rotation_center, const float3& linear_velocity, const float3& rotational_velocity) {
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->voxelize_mesh_on_device(mesh, flag, rotation_center, linear_velocity, rotational_velocity);
#ifdef MOVING_BOUNDARIES
	if((flag&(TYPE_S|TYPE_E))==TYPE_S&&(length(linear_velocity)>0.0f||length(rotational_velocity)>0.0f)) update_moving_boundaries();
#endif
	if(!initialized) { flags.read_from_device(); u.read_from_device(); } // so that initialize() does not overwrite the result with the host copies
}
void LBM::unvoxelize_mesh_on_device(const Mesh* mesh, const uchar flag) {
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->enqueue_unvoxelize_mesh_on_device(mesh, flag);
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->finish_queue();
}
void LBM::voxelize_stl(const string& path, const float3& center, const float3x3& rotation, const float size, const uchar flag) {
	const Mesh* mesh = read_stl(path, this->size(), center, rotation, size);
	flags.write_to_device();
	voxelize_mesh_on_device(mesh, flag);
	delete mesh;
	flags.read_from_device();
}

void LBM::initialize() {
#ifndef BENCHMARK
	sanity_checks_initialization();
#endif
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->rho.enqueue_write_to_device();
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->u.enqueue_write_to_device();
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->flags.enqueue_write_to_device();
#ifdef FORCE_FIELD
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->F.enqueue_write_to_device();
	if(get_D()>1u) communicate_F();
#endif
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->increment_time_step(); // halo exchange during initialisation runs at an odd step
	if(get_D()>1u) communicate_rho_u_flags();
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->enqueue_initialize();
	if(get_D()>1u) { communicate_rho_u_flags(); communicate_fi(); }
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->finish_queue();
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->reset_time_step();
	initialized = true;
}
void LBM::do_time_step() {
	if(fused_halo) { // the kernel delivers the y/z halo rows itself: one rendezvous, then only the x faces remain (see fx3d_stream_collide_fused)
		for(uint d=0u; d<get_D(); d++) {
			void* table[9] = { nullptr };
			for(int dz=-1; dz<=1; dz++) for(int dy=-1; dy<=1; dy++) if((dy==0||Dy>1u) && (dz==0||Dz>1u)) table[(dy+1)+3*(dz+1)] = lbm_domain[neighbour_yz(d, dy, dz)]->get_lattice().fi;
			lbm_domain[d]->enqueue_stream_collide_fused(table);
		}
		rendezvous();
		if(Dx>1u) communicate_field(FIELD_FI, 1u);
		for(uint d=0u; d<get_D(); d++) lbm_domain[d]->increment_time_step();
		return;
	}
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->enqueue_stream_collide();
	if(get_D()>1u) communicate_fi();
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->increment_time_step();
}
void LBM::run(const ulong steps, const ulong total_steps) {
	info.append(steps, total_steps, get_t());
	if(!initialized) {
		initialize();
		info.print_initialize(this);
	}
	// the device runs ahead of the host: steps are enqueued in chunks and timed per chunk, not per step
	const ulong chunk = 64ull;
	ulong done = 0ull;
	while(done<steps) {
		const ulong n = min(chunk, steps-done);
		Clock clock;
		if(get_D()==1u) { lbm_domain[0]->enqueue_run_steps(n); lbm_domain[0]->increment_time_step(n); }
		else for(ulong i=0ull; i<n; i++) do_time_step();
		for(uint d=0u; d<get_D(); d++) lbm_domain[d]->finish_queue();
		const double dt = clock.stop()/(double)n;
		for(ulong i=0ull; i<n; i++) info.update(dt);
		done += n;
	}
	if(get_D()>1u) for(uint d=0u; d<get_D(); d++) lbm_domain[d]->check_rendezvous();
}
void LBM::update_fields() {
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->enqueue_update_fields();
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->finish_queue();
}
void LBM::reset() { initialized = false; }

void LBM::write_status(const string& path) {
	const string filename = (path=="" ? get_exe_path() : path)+"status.txt";
	std::FILE* f = std::fopen(filename.c_str(), "w");
	if(!f) { print_warning("File \""+filename+"\" could not be written."); return; }
	std::fprintf(f, "Grid Resolution = (%u, %u, %u)\nGrid Domains = (%u, %u, %u)\nLBM type = D3Q%u %s\n", Nx, Ny, Nz, Dx, Dy, Dz, get_velocity_set(), info.collision.c_str());
	std::fprintf(f, "Memory Usage = CPU %u MB, GPU %ux %u MB\n", info.cpu_mem_required, get_D(), info.gpu_mem_required);
	std::fprintf(f, "Time Steps = %llu\nKinematic Viscosity = %.8f\nRelaxation Time = %.8f\nMaximum Reynolds Number = %.8f\n", (unsigned long long)get_t(), get_nu(), get_tau(), get_Re_max());
#ifdef VOLUME_FORCE
	std::fprintf(f, "Volume Force = (%.8f, %.8f, %.8f)\n", get_fx(), get_fy(), get_fz());
#endif
	std::fclose(f);
}
