// units.hpp -- the few lattice-unit helpers the hot-path fixture scenes and the VTK writer call through the global `units`
// object (the reference's Units class, FluidX3D v3.7 src/units.hpp, is OUT OF SCOPE as a component: SURVEY 2.1 #11).
// Scenes that need full SI conversion keep using the reference's header; only the member names below are shared.
#pragma once
#include "utilities.hpp"

struct Units {
	// SI size of one lattice length / mass / time unit; the fixtures run in lattice units, so all three stay 1
	float metres_per_cell = 1.0f, kilograms_per_unit = 1.0f, seconds_per_step = 1.0f;

	// scale factors the VTK writer puts into the file header (lbm.hpp write_vtk)
	float si_x(const float cells) const { return cells*metres_per_cell; }
	float si_u(const float lattice_speed) const { return lattice_speed*(metres_per_cell/seconds_per_step); }
	float si_rho(const float lattice_density) const { return lattice_density*(kilograms_per_unit/(metres_per_cell*metres_per_cell*metres_per_cell)); }

	// viscosity from the relaxation time: tau = 3 nu + 1/2
	float nu_from_tau(const float tau) const { return (tau-0.5f)*(1.0f/3.0f); }
	// viscosity that gives Reynolds number Re for a flow of speed `speed` past a body of `length` cells
	float nu_from_Re(const float Re, const float length, const float speed) const { return length*speed/Re; }
	// body force that drives a cylindrical Poiseuille flow of radius R to the centre-line speed u_max: u_max = f R^2 / (4 rho nu)
	float f_from_u_Poiseuille_3D(const float u_max, const float density, const float viscosity, const float R) const { return 4.0f*density*viscosity*u_max/(R*R); }
};
extern Units units; // defined in lbm.cpp
