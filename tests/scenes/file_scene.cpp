// file_scene.cpp -- TEST INFRASTRUCTURE ONLY. One main_setup() written against the reference's public host API (class LBM,
// Memory_Container, run(), read_from_device(), update_moving_boundaries(); FluidX3D v3.7 src/lbm.hpp:208-611) and compiled twice:
//   * in place of the reference's src/setup.cpp (oracle/ref/build_opencl_ref.py), so that the UNMODIFIED reference program runs
//     its own OpenCL implementation beside this repository's CUDA path on the same B200 (SURVEY 8d "Reference beside it (i)")
//   * against this repository's C++ host surface fluidx3d_b200/host/ (tests/test_host_scene.py): the same scene source is the
//     drop-in proof, and its output is compared with the oracle bit for bit.
// Everything is steered by environment variables, because the reference's command line carries device IDs (src/lbm.cpp:658):
//   FX3D_REF_IN     input file: uint32 Nx,Ny,Nz, then rho[N], ux[N], uy[N], uz[N] (float32) and flags[N] (uint8), n = x+(y+z*Ny)*Nx
//                   (absent: the constructor defaults rho=1, u=0, flags=0 and FX3D_REF_N="Nx,Ny,Nz")
//   FX3D_REF_D      "Dx,Dy,Dz" (default 1,1,1); FX3D_REF_NU (default 1.0); FX3D_REF_F "fx,fy,fz" (VOLUME_FORCE builds)
//   FX3D_REF_STEPS  time steps (default 0); FX3D_REF_OUT output file: rho, ux, uy, uz (float32) and flags (uint8) after the steps
//   FX3D_REF_BENCH  if set: after FX3D_REF_STEPS warm-up steps time this many steps with the host clock and print MLUPs/s
//   FX3D_REF_MB     "z,uy": MOVING_BOUNDARIES builds: after half the steps set u.y of the TYPE_S cells in plane z to uy and call
//                   update_moving_boundaries() (exercises src/lbm.cpp:1018-1027)
//   FX3D_REF_STL    "path,size": before initialisation voxelise the binary STL file with lbm.voxelize_stl(path, lbm.center(), R, size, TYPE_S|TYPE_X),
//                   R = rotation by atan2(0.6,0.8) about x (src/lbm.cpp:1130-1135: read_stl, flags to the device, kernel voxelize_mesh, flags back)
//   FX3D_REF_FF     "1": FORCE_FIELD builds: after the steps append to the output the force field F (x, y, z planes, after update_force_field) and 9
//                   floats: object_force, object_center_of_mass and object_torque (about the box centre) of the cells flagged TYPE_S|TYPE_X
#include "setup.hpp"
#include <cstdio>
#include <cstdlib>
#include <chrono>

static string env_or(const char* name, const string& fallback) { const char* v = getenv(name); return v ? string(v) : fallback; }
static vector<float> numbers(const string& s) { vector<float> out; size_t p = 0u; while(p<s.size()) { size_t q = s.find(',', p); if(q==string::npos) q = s.size(); out.push_back((float)atof(s.substr(p, q-p).c_str())); p = q+1u; } return out; }

void main_setup() {
	const vector<float> D = numbers(env_or("FX3D_REF_D", "1,1,1")), F = numbers(env_or("FX3D_REF_F", "0,0,0"));
	const float nu = (float)atof(env_or("FX3D_REF_NU", "1.0").c_str());
	const ulong steps = (ulong)atoll(env_or("FX3D_REF_STEPS", "0").c_str()), bench = (ulong)atoll(env_or("FX3D_REF_BENCH", "0").c_str());
	const string in = env_or("FX3D_REF_IN", ""), out = env_or("FX3D_REF_OUT", "");
	uint Nx = 0u, Ny = 0u, Nz = 0u;
	FILE* fin = nullptr;
	if(in!="") {
		fin = fopen(in.c_str(), "rb");
		if(!fin) { print_error("cannot open "+in); }
		uint h[3];
		if(fread(h, 4, 3, fin)!=3u) print_error("short header");
		Nx = h[0]; Ny = h[1]; Nz = h[2];
	} else {
		const vector<float> N = numbers(env_or("FX3D_REF_N", "64,64,64"));
		Nx = (uint)N[0]; Ny = (uint)N[1]; Nz = (uint)N[2];
	}
#ifdef VOLUME_FORCE
	LBM lbm(Nx, Ny, Nz, (uint)D[0], (uint)D[1], (uint)D[2], nu, F[0], F[1], F[2]);
#else
	LBM lbm(Nx, Ny, Nz, (uint)D[0], (uint)D[1], (uint)D[2], nu);
#endif
	const ulong N = lbm.get_N();
	const string stl = env_or("FX3D_REF_STL", "");
	if(stl!="") { // geometry first, as the reference's scenes do: voxelize_stl() before initialisation reads flags AND u back from the device (src/lbm.cpp:1084-1087)
		const size_t comma = stl.find(',');
		const float3x3 R(1.0f, 0.0f, 0.0f, 0.0f, 0.8f, -0.6f, 0.0f, 0.6f, 0.8f);
		lbm.voxelize_stl(stl.substr(0u, comma), lbm.center(), R, (float)atof(stl.substr(comma+1u).c_str()), TYPE_S|TYPE_X);
	}
	if(fin) { // the input fields go to every cell outside the voxelised body
		vector<float> buf(N);
		if(fread(buf.data(), 4, N, fin)!=N) print_error("short rho"); for(ulong n=0ull; n<N; n++) if(!(lbm.flags[n]&TYPE_X)) lbm.rho[n] = buf[n];
		if(fread(buf.data(), 4, N, fin)!=N) print_error("short ux");  for(ulong n=0ull; n<N; n++) if(!(lbm.flags[n]&TYPE_X)) lbm.u.x[n] = buf[n];
		if(fread(buf.data(), 4, N, fin)!=N) print_error("short uy");  for(ulong n=0ull; n<N; n++) if(!(lbm.flags[n]&TYPE_X)) lbm.u.y[n] = buf[n];
		if(fread(buf.data(), 4, N, fin)!=N) print_error("short uz");  for(ulong n=0ull; n<N; n++) if(!(lbm.flags[n]&TYPE_X)) lbm.u.z[n] = buf[n];
		vector<uchar> fl(N);
		if(fread(fl.data(), 1, N, fin)!=N) print_error("short flags"); for(ulong n=0ull; n<N; n++) if(!(lbm.flags[n]&TYPE_X)) lbm.flags[n] = fl[n];
		fclose(fin);
	}
	lbm.run(0u); // initialize
#ifdef MOVING_BOUNDARIES
	const string mb = env_or("FX3D_REF_MB", "");
	if(mb!="") {
		const vector<float> M = numbers(mb);
		lbm.run(steps/2ull);
		lbm.u.read_from_device(); lbm.flags.read_from_device(); // (update_fields overwrites u of fluid cells only with what the device holds)
		const uint z = (uint)M[0];
		for(uint y=0u; y<Ny; y++) for(uint x=0u; x<Nx; x++) { const ulong n = lbm.index(x, y, z); if((lbm.flags[n]&0x03)==TYPE_S) lbm.u.y[n] = M[1]; }
		lbm.u.write_to_device();
		lbm.update_moving_boundaries();
		lbm.run(steps-steps/2ull);
	} else
#endif
	lbm.run(steps);
	if(bench>0ull) {
		const auto t0 = std::chrono::steady_clock::now();
		lbm.run(bench);
		const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
		printf("\nFX3D_REF_RESULT mlups=%.1f steps=%llu seconds=%.6f N=%llu\n", (double)N*(double)bench/dt*1E-6, (unsigned long long)bench, dt, (unsigned long long)N);
	}
	if(out!="") {
		lbm.rho.read_from_device(); lbm.u.read_from_device(); lbm.flags.read_from_device();
		FILE* fo = fopen(out.c_str(), "wb");
		if(!fo) print_error("cannot open "+out);
		vector<float> buf(N);
		for(ulong n=0ull; n<N; n++) buf[n] = lbm.rho[n]; fwrite(buf.data(), 4, N, fo);
		for(ulong n=0ull; n<N; n++) buf[n] = lbm.u.x[n]; fwrite(buf.data(), 4, N, fo);
		for(ulong n=0ull; n<N; n++) buf[n] = lbm.u.y[n]; fwrite(buf.data(), 4, N, fo);
		for(ulong n=0ull; n<N; n++) buf[n] = lbm.u.z[n]; fwrite(buf.data(), 4, N, fo);
		vector<uchar> fl(N);
		for(ulong n=0ull; n<N; n++) fl[n] = lbm.flags[n]; fwrite(fl.data(), 1, N, fo);
#ifdef FORCE_FIELD
		if(env_or("FX3D_REF_FF", "")!="") {
			lbm.update_force_field();
			lbm.F.read_from_device();
			for(ulong n=0ull; n<N; n++) buf[n] = lbm.F.x[n]; fwrite(buf.data(), 4, N, fo);
			for(ulong n=0ull; n<N; n++) buf[n] = lbm.F.y[n]; fwrite(buf.data(), 4, N, fo);
			for(ulong n=0ull; n<N; n++) buf[n] = lbm.F.z[n]; fwrite(buf.data(), 4, N, fo);
			const float3 force = lbm.object_force(TYPE_S|TYPE_X), com = lbm.object_center_of_mass(TYPE_S|TYPE_X), torque = lbm.object_torque(lbm.center(), TYPE_S|TYPE_X);
			const float sums[9] = { force.x, force.y, force.z, com.x, com.y, com.z, torque.x, torque.y, torque.z };
			fwrite(sums, 4, 9, fo);
		}
#endif
		fclose(fo);
	}
	fflush(stdout);
}
