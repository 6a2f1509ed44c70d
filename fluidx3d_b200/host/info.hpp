// info.hpp -- run-time metrics and console table (the reference's Info singleton, FluidX3D v3.7 src/info.hpp, src/info.cpp).
// The smoothed step time defines the MLUPs/s the BENCHMARK scene reports: MLUPs = N*1e-6/runtime_lbm_timestep_smooth.
#pragma once
#include "utilities.hpp"
#include <mutex>

class LBM;
struct Info {
	LBM* lbm = nullptr;
	double runtime_lbm = 0.0, runtime_total = 0.0, runtime_total_last = 0.0;
	double runtime_lbm_timestep_last = 1.0, runtime_lbm_timestep_smooth = 1.0;
	Clock clock;
	ulong steps = max_ulong, steps_last = 0ull;
	uint cpu_mem_required = 0u, gpu_mem_required = 0u; // MB
	string collision = "";
	std::mutex allow_printing;
	void append(const ulong steps, const ulong total_steps, const ulong t);
	void update(const double dt); // one LBM step took dt seconds
	double time() const;          // elapsed, or estimated remaining time when the step count is known
	void print_logo() const;
	void print_initialize(LBM* lbm);
	void print_update() const;
	void print_finalize();
};
extern Info info;
