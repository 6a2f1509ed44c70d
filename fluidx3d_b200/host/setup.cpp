// setup.cpp -- scene scripts on the hot path, written against the same API as the reference's src/setup.cpp scenes:
// BENCHMARK (:5-36), 3D Taylor-Green vortices (:40-59), Poiseuille flow validation (:84-144), plus the two walled/open
// boxes SURVEY.md section 8d defines for the extensions on the path (cavity with a TYPE_E lid, wind tunnel with a sphere).
// Exactly one main_setup() is compiled: BENCHMARK from defines.hpp, otherwise -DSCENE_TAYLOR_GREEN / -DSCENE_POISEUILLE /
// -DSCENE_CAVITY / -DSCENE_WINDTUNNEL.
#include "setup.hpp"

#ifdef BENCHMARK
#include "info.hpp"
void main_setup() { // benchmark; required extensions in defines.hpp: BENCHMARK, optionally FP16S or FP16C
	uint mlups = 0u; {
		const char* size = std::getenv("FX3D_BENCHMARK_SIZE"); // default 256 like the reference; e.g. 512 for the FP16S headline
		const uint n = size ? (uint)std::atoi(size) : 256u;
		const char* split = std::getenv("FX3D_BENCHMARK_DOMAINS"); // e.g. "2x2x2"
		uint Dx=1u, Dy=1u, Dz=1u;
		if(split) std::sscanf(split, "%ux%ux%u", &Dx, &Dy, &Dz);
		LBM lbm(n*Dx, n*Dy, n*Dz, Dx, Dy, Dz, 1.0f);
		for(uint i=0u; i<100u; i++) {
			lbm.run(10u, 100u*10u);
			mlups = max(mlups, to_uint((double)lbm.get_N()*1E-6/info.runtime_lbm_timestep_smooth));
		}
	} // lbm goes out of scope here and frees its memory
	print_info("Peak MLUPs/s = "+to_string(mlups));
}
#endif // BENCHMARK

#ifdef SCENE_TAYLOR_GREEN
void main_setup() { // 3D Taylor-Green vortices in a periodic box
	LBM lbm(128u, 128u, 128u, 1u, 1u, 1u, 0.01f);
	const uint Nx=lbm.get_Nx(), Ny=lbm.get_Ny(), Nz=lbm.get_Nz();
	parallel_for(lbm.get_N(), [&](ulong n) { uint x=0u, y=0u, z=0u; lbm.coordinates(n, x, y, z);
		const float A = 0.25f;
		const float a=(float)Nx, b=(float)Ny, c=(float)Nz;
		const float fx = (float)x+0.5f-0.5f*(float)Nx, fy = (float)y+0.5f-0.5f*(float)Ny, fz = (float)z+0.5f-0.5f*(float)Nz;
		lbm.u.x[n] =  A*cosf(2.0f*pif*fx/a)*sinf(2.0f*pif*fy/b)*sinf(2.0f*pif*fz/c);
		lbm.u.y[n] = -A*sinf(2.0f*pif*fx/a)*cosf(2.0f*pif*fy/b)*sinf(2.0f*pif*fz/c);
		lbm.u.z[n] =  A*sinf(2.0f*pif*fx/a)*sinf(2.0f*pif*fy/b)*cosf(2.0f*pif*fz/c);
		lbm.rho[n] = 1.0f-sq(A)*3.0f/4.0f*(cosf(4.0f*pif*fx/a)+cosf(4.0f*pif*fy/b));
	});
	lbm.run(1000u);
	lbm.u.read_from_device();
	println("\nu.x at the box centre after 1000 steps: "+to_string(lbm.u.x[lbm.index(Nx/2u, Ny/2u, Nz/2u)], 8u)); // identity check between builds/devices
}
#endif // SCENE_TAYLOR_GREEN

#ifdef SCENE_POISEUILLE
void main_setup() { // Poiseuille flow validation; required extensions in defines.hpp: VOLUME_FORCE
	const uint R = 63u;          // channel radius
	const float umax = 0.1f;     // centre velocity
	const float tau = 1.0f;
	const float nu = units.nu_from_tau(tau);
	const uint H = 2u*(R+1u);
	LBM lbm(H, 4u, H, nu, 0.0f, units.f_from_u_Poiseuille_3D(umax, 1.0f, nu, (float)R), 0.0f);
	const uint Nx=lbm.get_Nx(), Ny=lbm.get_Ny(), Nz=lbm.get_Nz();
	parallel_for(lbm.get_N(), [&](ulong n) { uint x=0u, y=0u, z=0u; lbm.coordinates(n, x, y, z);
		if(!cylinder(x, y, z, lbm.center(), float3(0u, Ny, 0u), 0.5f*(float)min(Nx, Nz)-1.0f)) lbm.flags[n] = TYPE_S;
	});
	double error_min = max_double;
	while(true) {
		lbm.run(1000u);
		lbm.u.read_from_device();
		double error_dif=0.0, error_sum=0.0;
		const uint y = Ny/2u;
		for(uint x=0u; x<Nx; x++) for(uint z=0u; z<Nz; z++) {
			const ulong n = lbm.index(x, y, z);
			const double r = sqrt(sq((double)x+0.5-0.5*(double)Nx)+sq((double)z+0.5-0.5*(double)Nz));
			if(r<(double)R) {
				const double unum = sqrt(sq((double)lbm.u.x[n])+sq((double)lbm.u.y[n])+sq((double)lbm.u.z[n]));
				const double uref = umax*(sq((double)R)-sq(r))/sq((double)R);
				error_dif += sq(unum-uref); error_sum += sq(uref);
			}
		}
		const double error = sqrt(error_dif/error_sum);
		if(error>=error_min) { print_info("Poiseuille flow error converged after "+to_string(lbm.get_t())+" steps to "+to_string(100.0*error_min, 3u)+"%"); return; }
		error_min = fmin(error_min, error);
		print_info("Poiseuille flow error after t="+to_string(lbm.get_t())+" is "+to_string(100.0*error_min, 3u)+"%");
	}
}
#endif // SCENE_POISEUILLE

#ifdef SCENE_CAVITY
void main_setup() { // lid-driven cavity; required: MOVING_BOUNDARIES (the reference's formulation: a solid lid that moves) or EQUILIBRIUM_BOUNDARIES (TYPE_E lid carrying u)
	const uint L = 128u;
	const float Re = 1000.0f, u0 = 0.1f;
	LBM lbm(L, L, L, units.nu_from_Re(Re, (float)(L-2u), u0));
	const uint Nx=lbm.get_Nx(), Ny=lbm.get_Ny(), Nz=lbm.get_Nz();
	parallel_for(lbm.get_N(), [&](ulong n) { uint x=0u, y=0u, z=0u; lbm.coordinates(n, x, y, z);
#ifdef MOVING_BOUNDARIES
		if(z==Nz-1u) lbm.u.y[n] = u0;
		if(x==0u||x==Nx-1u||y==0u||y==Ny-1u||z==0u||z==Nz-1u) lbm.flags[n] = TYPE_S; // all non periodic
#else
		if(z==Nz-1u) { lbm.flags[n] = TYPE_E; lbm.u.y[n] = u0; }
		else if(x==0u||x==Nx-1u||y==0u||y==Ny-1u||z==0u) lbm.flags[n] = TYPE_S;
#endif
	});
	lbm.run(10000u);
	lbm.u.read_from_device();
	println("\nu.y at the cavity centre: "+to_string(lbm.u.y[lbm.index(Nx/2u, Ny/2u, Nz/2u)], 8u));
}
#endif // SCENE_CAVITY

#ifdef SCENE_WINDTUNNEL
void main_setup() { // sphere in a wind tunnel; required: D3Q27 or D3Q19, EQUILIBRIUM_BOUNDARIES, VOLUME_FORCE (TRT recommended)
	const uint3 N = uint3(256u, 512u, 256u);
	const float u0 = 0.075f, Re = 10000.0f;
	LBM lbm(N, units.nu_from_Re(Re, (float)N.x, u0), 0.0f, 1E-6f, 0.0f);
	const uint Nx=lbm.get_Nx(), Ny=lbm.get_Ny(), Nz=lbm.get_Nz();
	parallel_for(lbm.get_N(), [&](ulong n) { uint x=0u, y=0u, z=0u; lbm.coordinates(n, x, y, z);
		if(sphere(x, y, z, float3(0.5f*(float)Nx, 0.25f*(float)Ny, 0.5f*(float)Nz), 0.125f*(float)Nx)) lbm.flags[n] = TYPE_S;
		else lbm.u.y[n] = u0;
		if(x==0u||x==Nx-1u||y==0u||y==Ny-1u||z==0u||z==Nz-1u) lbm.flags[n] = TYPE_E;
	});
	lbm.run(2000u);
	lbm.u.read_from_device();
	println("\nu.y behind the sphere: "+to_string(lbm.u.y[lbm.index(Nx/2u, Ny/2u, Nz/2u)], 8u));
}
#endif // SCENE_WINDTUNNEL
