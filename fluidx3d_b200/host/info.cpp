#include "info.hpp"
#include "lbm.hpp"

Info info;

void Info::append(const ulong steps_, const ulong total_steps, const ulong t) {
	if(total_steps==max_ulong) {
		steps = steps_;
		steps_last = t;
		runtime_total_last = runtime_total;
		runtime_total = clock.stop();
	} else steps = total_steps;
}
void Info::update(const double dt) {
	runtime_lbm_timestep_last = dt;
	runtime_lbm_timestep_smooth = (dt+0.3)/(0.3/runtime_lbm_timestep_smooth+1.0); // the reference's smoothing (src/info.cpp:18)
	runtime_lbm += dt;
	runtime_total = clock.stop();
}
double Info::time() const {
	if(lbm==nullptr) return 0.0;
	if(steps==max_ulong) return runtime_total;
	const double done = (double)max(lbm->get_t()-steps_last, (ulong)1ull);
	return ((double)steps/done-1.0)*(runtime_total-runtime_total_last);
}
void Info::print_logo() const {
	println("fx3d-b200 -- lattice Boltzmann stream_collide + halo exchange for NVIDIA B200 behind the FluidX3D host API");
}
void Info::print_initialize(LBM* l) {
	std::lock_guard<std::mutex> lock(allow_printing);
	lbm = l;
#if defined(SRT)
	collision = "SRT";
#else
	collision = "TRT";
#endif
#if defined(FP16S)
	collision += " (FP32/FP16S)";
#elif defined(FP16C)
	collision += " (FP32/FP16C)";
#else
	collision += " (FP32/FP32)";
#endif
	cpu_mem_required = (uint)(lbm->get_N()*(ulong)bytes_per_cell_host()/1048576ull);
	gpu_mem_required = lbm->lbm_domain[0]->get_device().info.memory_used;
	// one summary line instead of the reference's console table (the table is OUT OF SCOPE; the metric is not)
	char b[512];
	std::snprintf(b, sizeof(b), "grid %ux%ux%u = %llu cells in %ux%ux%u domains | D3Q%u %s | nu %.8g tau %.8g | host %u MB, device %ux %u MB | steps %s",
		lbm->get_Nx(), lbm->get_Ny(), lbm->get_Nz(), (unsigned long long)lbm->get_N(), lbm->get_Dx(), lbm->get_Dy(), lbm->get_Dz(), lbm->get_velocity_set(), collision.c_str(),
		(double)lbm->get_nu(), (double)lbm->get_tau(), cpu_mem_required, lbm->get_D(), gpu_mem_required, steps==max_ulong ? "unbounded" : to_string(steps).c_str());
	println(b);
	clock.start();
}
void Info::print_update() const {
	if(lbm==nullptr) return;
	std::lock_guard<std::mutex> lock(const_cast<std::mutex&>(allow_printing));
	if(lbm==nullptr) return;
	const double dt = runtime_lbm_timestep_smooth; // MLUPs/s = N*1e-6/dt (src/info.cpp:109), GB/s = MLUPs * bytes per cell and step (src/lbm.cpp:52)
	char b[256];
	std::snprintf(b, sizeof(b), "\r%9u MLUPs/s %7u GB/s %8u steps/s   t = %llu   %s %.0f s   ", to_uint((double)lbm->get_N()*1E-6/dt),
		to_uint((double)lbm->get_N()*(double)bandwidth_bytes_per_cell_device()*1E-9/dt), to_uint(1.0/dt), (unsigned long long)lbm->get_t(), steps==max_ulong ? "elapsed" : "remaining", time());
	std::cout << b << std::flush;
}
void Info::print_finalize() {
	std::lock_guard<std::mutex> lock(allow_printing);
	lbm = nullptr;
	println("");
}
