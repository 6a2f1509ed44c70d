// units.hpp -- conversion between SI and lattice units, same interface as the reference's Units class
// (FluidX3D v3.7 src/units.hpp) for the members scenes on the hot path use. Each base unit is the size of one lattice
// unit in SI: [m], [kg], [s].
#pragma once
#include "utilities.hpp"

class Units {
	float unit_m = 1.0f, unit_kg = 1.0f, unit_s = 1.0f;
	void report() const { print_info("Unit Conversion: 1 cell = "+to_string(1000.0f*si_x(1.0f), 3u)+" mm, 1 s = "+to_string(t(1.0f))+" time steps"); }
public:
	// fix the three base units from one length, one velocity and one density given in both systems
	void set_m_kg_s(const float x, const float u, const float rho, const float si_x, const float si_u, const float si_rho) {
		unit_m = si_x/x;
		unit_kg = si_rho/rho*cb(unit_m);
		unit_s = u/si_u*unit_m;
		report();
	}
	void set_m_kg_s(const float m, const float kg, const float s) { unit_m = m; unit_kg = kg; unit_s = s; report(); }

	// SI -> lattice
	float x(const float si_x) const { return si_x/unit_m; }
	float m(const float si_m) const { return si_m/unit_kg; }
	ulong t(const float si_t) const { return (ulong)(si_t/unit_s+0.5f); }
	float frequency(const float si_frequency) const { return si_frequency*unit_s; }
	float u(const float si_u) const { return si_u*unit_s/unit_m; }
	float rho(const float si_rho) const { return si_rho*cb(unit_m)/unit_kg; }
	float nu(const float si_nu) const { return si_nu*unit_s/sq(unit_m); }
	float mu(const float si_mu) const { return si_mu*unit_s*unit_m/unit_kg; }
	float g(const float si_g) const { return si_g/unit_m*sq(unit_s); }
	float f(const float si_f) const { return si_f*sq(unit_m*unit_s)/unit_kg; }
	float f(const float si_rho, const float si_g) const { return si_rho*si_g*sq(unit_m*unit_s)/unit_kg; }
	float F(const float si_F) const { return si_F*sq(unit_s)/(unit_kg*unit_m); }
	// lattice -> SI
	float si_x(const uint x) const { return (float)x*unit_m; }
	float si_x(const float x) const { return x*unit_m; }
	float si_m(const float m) const { return m*unit_kg; }
	float si_t(const ulong t) const { return (float)t*unit_s; }
	float si_u(const float u) const { return u*unit_m/unit_s; }
	float si_rho(const float rho) const { return rho*unit_kg/cb(unit_m); }
	float si_p(const float p) const { return p*unit_kg/(unit_m*sq(unit_s)); }
	float si_nu(const float nu) const { return nu*sq(unit_m)/unit_s; }
	float si_f(const float f) const { return f*unit_kg/sq(unit_m*unit_s); }
	float si_F(const float F) const { return F*unit_kg*unit_m/sq(unit_s); }
	// dimensionless numbers and lattice-unit relations
	float Re(const float x, const float u, const float nu) const { return x*u/nu; }
	float Ma(const float u) const { return u/0.57735027f; }
	float p_from_rho(const float rho) const { return (rho-1.0f)/3.0f; }
	float rho_from_p(const float p) const { return 1.0f+3.0f*p; }
	float nu_from_mu(const float mu, const float rho) const { return mu/rho; }
	float nu_from_tau(const float tau) const { return (tau-0.5f)/3.0f; }
	float nu_from_Re(const float Re, const float x, const float u) const { return x*u/Re; }
	float u_from_Re(const float Re, const float x, const float nu) const { return Re*nu/x; }
	float u_from_Ma(const float Ma) const { return 0.57735027f*Ma; }
	float f_from_g(const float g, const float rho) const { return rho*g; }
	float u_from_f_Poiseuille_2D(const float f, const float rho, const float nu, const float R) const { return f*sq(R)/(2.0f*rho*nu); }
	float u_from_f_Poiseuille_3D(const float f, const float rho, const float nu, const float R) const { return f*sq(R)/(4.0f*rho*nu); }
	float f_from_u_Poiseuille_2D(const float u, const float rho, const float nu, const float R) const { return 2.0f*u*rho*nu/sq(R); }
	float f_from_u_Poiseuille_3D(const float u, const float rho, const float nu, const float R) const { return 4.0f*u*rho*nu/sq(R); }
};
extern Units units; // defined in lbm.cpp
