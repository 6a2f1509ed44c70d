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
	println(".-----------------------------------------------------------------------------.");
	println("|  fx3d-b200: lattice Boltzmann stream_collide + halo exchange for NVIDIA B200 |");
	println("|  behind the FluidX3D host API (LBM / LBM_Domain / Memory<T>)                 |");
	println("|-----------------------------------------------------------------------------|");
}
static string hms(const double seconds) {
	const ulong s = (ulong)max(seconds, 0.0);
	char b[64];
	std::snprintf(b, sizeof(b), "%luh %02lum %02lus", (unsigned long)(s/3600ull), (unsigned long)((s/60ull)%60ull), (unsigned long)(s%60ull));
	return b;
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
	const float Re = lbm->get_Re_max();
	println("|-----------------.-----------------------------------------------------------|");
	println("| Grid Resolution | "+alignr(57u, to_string(lbm->get_Nx())+" x "+to_string(lbm->get_Ny())+" x "+to_string(lbm->get_Nz())+" = "+to_string(lbm->get_N()))+" |");
	println("| Grid Domains    | "+alignr(57u, to_string(lbm->get_Dx())+" x "+to_string(lbm->get_Dy())+" x "+to_string(lbm->get_Dz())+" = "+to_string(lbm->get_D()))+" |");
	println("| LBM Type        | "+alignr(57u, "D3Q"+to_string(lbm->get_velocity_set())+" "+collision)+" |");
	println("| Memory Usage    | "+alignr(54u, "CPU "+to_string(cpu_mem_required)+" MB, GPU "+to_string(lbm->get_D())+"x "+to_string(gpu_mem_required))+" MB |");
	println("| Time Steps      | "+alignr(57u, steps==max_ulong ? string("infinite") : to_string(steps))+" |");
	println("| Kin. Viscosity  | "+alignr(57u, to_string(lbm->get_nu(), 8u))+" |");
	println("| Relaxation Time | "+alignr(57u, to_string(lbm->get_tau(), 8u))+" |");
	println("| Reynolds Number | "+alignr(57u, "Re < "+(Re>=100.0f ? to_string(to_uint(Re)) : to_string(Re, 6u)))+" |");
#ifdef VOLUME_FORCE
	println("| Volume Force    | "+alignr(57u, alignr(15u, to_string(lbm->get_fx(), 8u))+","+alignr(15u, to_string(lbm->get_fy(), 8u))+","+alignr(15u, to_string(lbm->get_fz(), 8u)))+" |");
#endif
	println("|---------.-------'-----.-----------.-------------------.---------------------|");
	println("| MLUPs   | Bandwidth   | Steps/s   | Current Step      | "+string(steps==max_ulong ? "Elapsed Time  " : "Time Remaining")+"      |");
	clock.start();
}
void Info::print_update() const {
	if(lbm==nullptr) return;
	std::lock_guard<std::mutex> lock(const_cast<std::mutex&>(allow_printing));
	if(lbm==nullptr) return;
	const double dt = runtime_lbm_timestep_smooth;
	std::cout << "\r|" << alignr(8u, to_string(to_uint((double)lbm->get_N()*1E-6/dt))) << " |"
		<< alignr(7u, to_string(to_uint((double)lbm->get_N()*(double)bandwidth_bytes_per_cell_device()*1E-9/dt))) << " GB/s |"
		<< alignr(10u, to_string(to_uint(1.0/dt))) << " | " << alignr(17u, to_string(lbm->get_t())) << " | " << alignr(19u, hms(time())) << " |" << std::flush;
}
void Info::print_finalize() {
	std::lock_guard<std::mutex> lock(allow_printing);
	lbm = nullptr;
	println("\n|---------'-------------'-----------'-------------------'---------------------|");
}
