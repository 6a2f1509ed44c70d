/*
 * lbm_oracle.c -- CPU oracle (plain C restatement) of the FluidX3D LBM hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see lbm_oracle.h. Never linked into, or called from, the product.
 *
 * Compile with strict IEEE semantics: -O2 -ffp-contract=off -fno-fast-math (oracle/Makefile does).
 * Every fma() below is an explicit fma() in the reference; every other operation is a separately
 * rounded binary32 operation evaluated left to right exactly as the reference expression is written.
 *
 * Reference citations are "file:line" into the FluidX3D v3.7 tree (src/...).
 */
#include "lbm_oracle.h"
#include <math.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define QMAX 27u
#define TYPE_S 0x01u
#define TYPE_E 0x02u
#define TYPE_BO 0x03u  /* src/lbm.cpp:393-408 */
#define TYPE_MS 0x03u  /* cell next to a moving solid boundary */
#define TYPE_T 0x04u

static int g_threads = 0;
void orc_set_threads(int n) { g_threads = n; }
int orc_get_threads(void) {
#ifdef _OPENMP
	return g_threads>0 ? g_threads : omp_get_max_threads();
#else
	return 1;
#endif
}
#ifdef _OPENMP
#define ORC_PARALLEL_FOR _Pragma("omp parallel for schedule(static) num_threads(orc_get_threads())")
#else
#define ORC_PARALLEL_FOR
#endif

static inline uint32_t bits_of(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
static inline float float_of(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }

/* ---- velocity set: src/kernel.cpp:864-881 (c), :882-903 (w), direction numbering also :933-956 ---- */
static const int8_t EX[QMAX] = { 0, 1,-1, 0, 0, 0, 0, 1,-1, 1,-1, 0, 0, 1,-1, 1,-1, 0, 0, 1,-1, 1,-1, 1,-1,-1, 1 };
static const int8_t EY[QMAX] = { 0, 0, 0, 1,-1, 0, 0, 1,-1, 0, 0, 1,-1,-1, 1, 0, 0, 1,-1, 1,-1, 1,-1,-1, 1, 1,-1 };
static const int8_t EZ[QMAX] = { 0, 0, 0, 0, 0, 1,-1, 0, 0, 1,-1, 1,-1, 0, 0,-1, 1,-1, 1, 1,-1,-1, 1, 1,-1, 1,-1 };

typedef struct { float w0, ws, we, wc; } weights_t;
static weights_t weights_of(uint32_t Q) { /* src/lbm.cpp:376-384: float constant expressions */
	weights_t k;
	if(Q==19u) { k.w0 = 1.0f/3.0f;   k.ws = 1.0f/18.0f; k.we = 1.0f/36.0f; k.wc = 0.0f; }
	else       { k.w0 = 1.0f/3.375f; k.ws = 1.0f/13.5f; k.we = 1.0f/54.0f; k.wc = 1.0f/216.0f; }
	return k;
}
static inline float weight_of(const weights_t* k, uint32_t i) { /* src/kernel.cpp:882-903 */
	return i==0u ? k->w0 : i<7u ? k->ws : i<19u ? k->we : k->wc;
}

/* ---- storage codecs ---- */
uint16_t orc_fp16s_encode(float x) { /* src/lbm.cpp:414: vstore_half_rte(x*32768.0f) */
	const _Float16 h = (_Float16)(x*32768.0f); /* IEEE binary16, round to nearest even, overflow -> inf */
	uint16_t r; memcpy(&r, &h, 2); return r;
}
float orc_fp16s_decode(uint16_t h) { /* src/lbm.cpp:413: vload_half(...)*3.0517578E-5f */
	_Float16 v; memcpy(&v, &h, 2);
	return (float)v*3.0517578E-5f;
}
float orc_fp16c_decode(uint16_t x) { /* src/kernel.cpp:848-853 */
	const uint32_t sign = ((uint32_t)x&0x8000u)<<16;
	const uint32_t e = ((uint32_t)x&0x7800u)>>11;
	const uint32_t m = ((uint32_t)x&0x07FFu)<<12;
	if(e!=0u) return float_of(sign|((e+112u)<<23)|m); /* normalised */
	if(m!=0u) { /* denormalised: exponent of (float)m locates the leading one */
		const uint32_t v = bits_of((float)m)>>23;
		return float_of(sign|((v-37u)<<23)|((m<<(150u-v))&0x007FF000u));
	}
	return float_of(sign);
}
uint16_t orc_fp16c_encode(float x) { /* src/kernel.cpp:854-859 (device version: no saturation term) */
	const uint32_t b = bits_of(x)+0x00000800u;
	const uint32_t e = (b&0x7F800000u)>>23;
	const uint32_t m = b&0x007FFFFFu;
	uint32_t r = (b&0x80000000u)>>16;
	if(e>112u) r |= (((e-112u)<<11)&0x7800u)|(m>>12);
	else if(e>100u) r |= (((0x007FF800u+m)>>(124u-e))+1u)>>1;
	return (uint16_t)r;
}

static inline float load_ddf(const orc_grid* g, const void* fi, uint64_t idx) { /* load(p,o) macro, src/lbm.cpp:410-425 */
	switch(g->storage) {
		case ORC_FP16S: return orc_fp16s_decode(((const uint16_t*)fi)[idx]);
		case ORC_FP16C: return orc_fp16c_decode(((const uint16_t*)fi)[idx]);
		default: return ((const float*)fi)[idx];
	}
}
static inline void store_ddf(const orc_grid* g, void* fi, uint64_t idx, float v) { /* store(p,o,x) macro */
	switch(g->storage) {
		case ORC_FP16S: ((uint16_t*)fi)[idx] = orc_fp16s_encode(v); break;
		case ORC_FP16C: ((uint16_t*)fi)[idx] = orc_fp16c_encode(v); break;
		default: ((float*)fi)[idx] = v;
	}
}

/* ---- decimal round trip of def_w: src/utilities.hpp:2599-2630 (split_float), :2669-2676, :2745-2754 ---- */
int orc_float_to_string(float x, char* out, int cap) {
	char sign[2] = { 0, 0 };
	if(x<0.0f) { sign[0] = '-'; x = -x; }
	if(isnan(x)) return snprintf(out, (size_t)cap, "%sNaN", sign);
	if(isinf(x)) return snprintf(out, (size_t)cap, "%sInf", sign);
	int exponent = 0;
	if(x>=10.0f) {
		if(x>=1E32f) { x *= 1E-32f; exponent += 32; }
		if(x>=1E16f) { x *= 1E-16f; exponent += 16; }
		if(x>= 1E8f) { x *=  1E-8f; exponent +=  8; }
		if(x>= 1E4f) { x *=  1E-4f; exponent +=  4; }
		if(x>= 1E2f) { x *=  1E-2f; exponent +=  2; }
		if(x>= 1E1f) { x *=  1E-1f; exponent +=  1; }
	}
	if(x>0.0f && x<=1.0f) {
		if(x<1E-31f) { x *=  1E32f; exponent -= 32; }
		if(x<1E-15f) { x *=  1E16f; exponent -= 16; }
		if(x< 1E-7f) { x *=   1E8f; exponent -=  8; }
		if(x< 1E-3f) { x *=   1E4f; exponent -=  4; }
		if(x< 1E-1f) { x *=   1E2f; exponent -=  2; }
		if(x<  1E0f) { x *=   1E1f; exponent -=  1; }
	}
	uint32_t integral = (uint32_t)x;
	const float remainder = (x-(float)integral)*1E8f; /* 8 decimal digits, float arithmetic */
	uint32_t decimal = (uint32_t)remainder;
	if(remainder-(float)decimal>=0.5f) {
		decimal++;
		if(decimal>=100000000u) {
			decimal = 0u;
			integral++;
			if(integral>=10u) { integral = 1u; exponent++; }
		}
	}
	if(exponent!=0) return snprintf(out, (size_t)cap, "%s%u.%08uE%d", sign, integral, decimal, exponent);
	return snprintf(out, (size_t)cap, "%s%u.%08u", sign, integral, decimal);
}
float orc_w_from_nu(float nu) { /* src/lbm.hpp:148 (tau = 3*nu+0.5), src/lbm.cpp:367 ("#define def_w "+to_string(1.0f/tau)+"f") */
	const float tau = 3.0f*nu+0.5f;
	char s[64];
	orc_float_to_string(1.0f/tau, s, (int)sizeof(s));
	return strtof(s, NULL); /* the OpenCL C compiler parses the decimal literal, correctly rounded */
}

/* ---- indexing: src/kernel.cpp:820-826, :843-846, :861-863, :904-958 ---- */
typedef struct { uint32_t x, y, z; } xyz_t;
static inline xyz_t coords_of(const orc_grid* g, uint64_t n) {
	xyz_t c;
	const uint64_t plane = (uint64_t)g->Nx*g->Ny, t = n%plane;
	c.x = (uint32_t)(t%g->Nx); c.y = (uint32_t)(t/g->Nx); c.z = (uint32_t)(n/plane);
	return c;
}
static inline uint64_t index_of(const orc_grid* g, uint32_t x, uint32_t y, uint32_t z) {
	return (uint64_t)x+((uint64_t)y+(uint64_t)z*g->Ny)*g->Nx;
}
static inline uint64_t cells_of(const orc_grid* g) { return (uint64_t)g->Nx*g->Ny*g->Nz; }
static inline int halo_cell(const orc_grid* g, xyz_t c) {
	return (g->Dx>1u&&(c.x==0u||c.x>=g->Nx-1u))||(g->Dy>1u&&(c.y==0u||c.y>=g->Ny-1u))||(g->Dz>1u&&(c.z==0u||c.z>=g->Nz-1u));
}
static inline uint64_t neighbour_of(const orc_grid* g, xyz_t c, uint32_t i) { /* periodic inside the local (halo-inclusive) box */
	const uint32_t x = (uint32_t)((c.x+g->Nx+(uint32_t)(int32_t)EX[i])%g->Nx);
	const uint32_t y = (uint32_t)((c.y+g->Ny+(uint32_t)(int32_t)EY[i])%g->Ny);
	const uint32_t z = (uint32_t)((c.z+g->Nz+(uint32_t)(int32_t)EZ[i])%g->Nz);
	return index_of(g, x, y, z);
}
static void neighbours_of(const orc_grid* g, uint64_t n, uint64_t* j) {
	const xyz_t c = coords_of(g, n);
	j[0] = n;
	for(uint32_t i=1u; i<g->Q; i++) j[i] = neighbour_of(g, c, i);
}
static inline uint64_t slot_index(const orc_grid* g, uint64_t n, uint32_t i) { return (uint64_t)i*cells_of(g)+n; }

/* ---- Esoteric-Pull addressing: src/kernel.cpp:1326-1339 ---- */
static void pull_ddfs(const orc_grid* g, uint64_t n, float* fhn, const void* fi, const uint64_t* j, uint64_t t) {
	const uint32_t odd = (uint32_t)(t&1ull);
	fhn[0] = load_ddf(g, fi, slot_index(g, n, 0u));
	for(uint32_t i=1u; i<g->Q; i+=2u) {
		fhn[i   ] = load_ddf(g, fi, slot_index(g, n   , odd ? i    : i+1u));
		fhn[i+1u] = load_ddf(g, fi, slot_index(g, j[i], odd ? i+1u : i   ));
	}
}
static void push_ddfs(const orc_grid* g, uint64_t n, const float* fhn, void* fi, const uint64_t* j, uint64_t t) {
	const uint32_t odd = (uint32_t)(t&1ull);
	store_ddf(g, fi, slot_index(g, n, 0u), fhn[0]);
	for(uint32_t i=1u; i<g->Q; i+=2u) {
		store_ddf(g, fi, slot_index(g, j[i], odd ? i+1u : i   ), fhn[i   ]);
		store_ddf(g, fi, slot_index(g, n   , odd ? i    : i+1u), fhn[i+1u]);
	}
}

/* ---- equilibrium: src/kernel.cpp:1004-1061 ---- */
static void equilibrium(const orc_grid* g, float rho, float ux, float uy, float uz, float* feq) {
	const weights_t k = weights_of(g->Q);
	const float rhom1 = rho-1.0f;
	const float c3 = -3.0f*(ux*ux+uy*uy+uz*uz); /* sq(ux)+sq(uy)+sq(uz), left to right */
	uz *= 3.0f; ux *= 3.0f; uy *= 3.0f;
	feq[0] = k.w0*fmaf(rho, 0.5f*c3, rhom1);
	const float comp[3] = { ux, uy, uz };
	for(uint32_t i=1u; i<g->Q; i+=2u) {
		/* projected velocity of the "+" member: first non-zero component (negated if the direction says so),
		   then the remaining ones added/subtracted in x,y,z order -- reproduces u0..u9 of :1033/:1045 */
		const int8_t e[3] = { EX[i], EY[i], EZ[i] };
		float uq = 0.0f; int first = 1;
		for(int a=0; a<3; a++) if(e[a]!=0) {
			if(first) { uq = e[a]>0 ? comp[a] : -comp[a]; first = 0; }
			else uq = e[a]>0 ? uq+comp[a] : uq-comp[a];
		}
		const float wq = weight_of(&k, i);
		const float rhoq = wq*rho, rhom1q = wq*rhom1;
		feq[i   ] = fmaf(rhoq, fmaf(0.5f, fmaf(uq, uq, c3),  uq), rhom1q);
		feq[i+1u] = fmaf(rhoq, fmaf(0.5f, fmaf(uq, uq, c3), -uq), rhom1q);
	}
}

/* ---- moments: src/kernel.cpp:1063-1088 ---- */
static void moments(const orc_grid* g, const float* f, float* rhon, float* uxn, float* uyn, float* uzn) {
	float rho = f[0];
	for(uint32_t i=1u; i<g->Q; i++) rho += f[i];
	rho += 1.0f;
	float mom[3];
	for(int a=0; a<3; a++) { /* alternating sums: for each pair with a component along axis a, positive member first */
		const int8_t* E = a==0 ? EX : a==1 ? EY : EZ;
		float s = 0.0f; int first = 1;
		for(uint32_t i=1u; i<g->Q; i+=2u) if(E[i]!=0) {
			const float fp = E[i]>0 ? f[i] : f[i+1u], fm = E[i]>0 ? f[i+1u] : f[i];
			if(first) { s = fp-fm; first = 0; } else { s = s+fp; s = s-fm; }
		}
		mom[a] = s;
	}
	*rhon = rho; *uxn = mom[0]/rho; *uyn = mom[1]/rho; *uzn = mom[2]/rho;
}

/* ---- Guo forcing: src/kernel.cpp:1090-1102 ---- */
static void forcing_terms(const orc_grid* g, float ux, float uy, float uz, float fx, float fy, float fz, float* Fin) {
	const weights_t k = weights_of(g->Q);
	const float uF = -0.33333334f*fmaf(ux, fx, fmaf(uy, fy, uz*fz));
	Fin[0] = 9.0f*k.w0*uF;
	for(uint32_t i=1u; i<g->Q; i++) {
		const float cx = (float)EX[i], cy = (float)EY[i], cz = (float)EZ[i];
		Fin[i] = 9.0f*weight_of(&k, i)*fmaf(cx*fx+cy*fy+cz*fz, cx*ux+cy*uy+cz*uz+0.33333334f, uF);
	}
}

static inline float clamp_c(float x) { const float c = 0.57735027f; return fminf(fmaxf(x, -c), c); } /* OpenCL clamp(), def_c src/lbm.cpp:366 */

/* is any neighbour of n a TYPE_S cell with non-zero velocity? (:1381-1385, :1442-1446) */
static int next_to_moving_solid(const orc_grid* g, const uint64_t* j, const float* u, const uint8_t* flags) {
	const uint64_t N = cells_of(g);
	int r = 0;
	for(uint32_t i=1u; i<g->Q; i++) r = r || ((flags[j[i]]&TYPE_BO)==TYPE_S && (u[j[i]]!=0.0f || u[N+j[i]]!=0.0f || u[2ull*N+j[i]]!=0.0f));
	return r;
}
/* ---- apply_moving_boundaries: src/kernel.cpp:1104-1113 (Dirichlet velocity boundary, rho_wall = 1) ---- */
static void apply_moving_boundaries(const orc_grid* g, float* fhn, const uint64_t* j, const float* u, const uint8_t* flags) {
	const uint64_t N = cells_of(g);
	const weights_t k = weights_of(g->Q);
	for(uint32_t i=1u; i<g->Q; i+=2u) {
		const float w6 = -6.0f*weight_of(&k, i);
		uint64_t ji = j[i+1u];
		if((flags[ji]&TYPE_BO)==TYPE_S) fhn[i   ] = fmaf(w6, (float)EX[i+1u]*u[ji]+(float)EY[i+1u]*u[N+ji]+(float)EZ[i+1u]*u[2ull*N+ji], fhn[i   ]);
		ji = j[i];
		if((flags[ji]&TYPE_BO)==TYPE_S) fhn[i+1u] = fmaf(w6, (float)EX[i   ]*u[ji]+(float)EY[i   ]*u[N+ji]+(float)EZ[i   ]*u[2ull*N+ji], fhn[i+1u]);
	}
}
/* ---- update_moving_boundaries: src/kernel.cpp:1432-1450 ---- */
void orc_update_moving_boundaries(const orc_grid* g, const float* u, uint8_t* flags) {
	const uint64_t N = cells_of(g);
	ORC_PARALLEL_FOR
	for(uint64_t n=0ull; n<N; n++) {
		const xyz_t c = coords_of(g, n);
		if(halo_cell(g, c)) continue;
		const uint8_t fn = flags[n], fb = fn&TYPE_BO;
		if(fb==TYPE_S || fb==TYPE_E || (fn&TYPE_T)) continue;
		uint64_t j[QMAX];
		neighbours_of(g, n, j);
		flags[n] = next_to_moving_solid(g, j, u, flags) ? (uint8_t)(fn|TYPE_MS) : (uint8_t)(fn&~TYPE_MS);
	}
}

/* ---- initialize: src/kernel.cpp:1358-1430 (non-SURFACE, non-TEMPERATURE build) ---- */
void orc_initialize(const orc_grid* g, void* fi, const float* rho, float* u, uint8_t* flags) {
	const uint64_t N = cells_of(g);
	const int mb = (g->features&ORC_MOVING_BOUNDARIES)!=0u;
	ORC_PARALLEL_FOR
	for(uint64_t n=0ull; n<N; n++) {
		const xyz_t c = coords_of(g, n);
		if(halo_cell(g, c)) continue;
		uint64_t j[QMAX]; float feq[QMAX];
		neighbours_of(g, n, j);
		const uint8_t fb = flags[n]&TYPE_BO;
		if(!mb) { if(fb==TYPE_S) { u[n] = 0.0f; u[N+n] = 0.0f; u[2ull*N+n] = 0.0f; } } /* :1376-1379 */
		else if(fb==TYPE_S) { /* :1374-1377: only solids enclosed by solids lose their velocity */
			int only_s = 1;
			for(uint32_t i=1u; i<g->Q; i++) only_s = only_s && (flags[j[i]]&TYPE_BO)==TYPE_S;
			if(only_s) { u[n] = 0.0f; u[N+n] = 0.0f; u[2ull*N+n] = 0.0f; }
		} else if(fb!=TYPE_E) flags[n] = next_to_moving_solid(g, j, u, flags) ? (uint8_t)(flags[n]|TYPE_MS) : (uint8_t)(flags[n]&~TYPE_MS); /* :1380-1386 */
		equilibrium(g, rho[n], u[n], u[N+n], u[2ull*N+n], feq);
		push_ddfs(g, n, feq, fi, j, 1ull); /* :1429: odd-step layout */
	}
}

/* shared front half of stream_collide / update_fields: load, moments (or preset), force shift, clamp */
static int cell_front(const orc_grid* g, const void* fi, const float* rho, const float* u, const uint8_t* flags, uint64_t n, uint64_t t,
	float* fxp, float* fyp, float* fzp, const float* F, int allow_preset, uint64_t* j, float* fhn, float* rhon, float* uxn, float* uyn, float* uzn, uint8_t* flag_bo) {
	const uint64_t N = cells_of(g);
	const xyz_t c = coords_of(g, n);
	if(halo_cell(g, c)) return 0;
	const uint8_t fb = flags[n]&TYPE_BO;
	if(fb==TYPE_S) return 0; /* :1469 / :1806 (TYPE_G never set without SURFACE) */
	*flag_bo = fb;
	neighbours_of(g, n, j);
	pull_ddfs(g, n, fhn, fi, j, t);
	if((g->features&ORC_MOVING_BOUNDARIES) && fb==TYPE_MS) apply_moving_boundaries(g, fhn, j, u, flags); /* :1477-1479 / :1813-1815 */
	if(allow_preset && (g->features&ORC_EQUILIBRIUM_BOUNDARIES) && fb==TYPE_E) { /* :1482-1493 */
		*rhon = rho[n]; *uxn = u[n]; *uyn = u[N+n]; *uzn = u[2ull*N+n];
	} else moments(g, fhn, rhon, uxn, uyn, uzn);
	float fx = *fxp, fy = *fyp, fz = *fzp; /* :1494 / :1819: the force starts as the constant volume force */
	if((g->features&ORC_FORCE_FIELD) && F) { fx += F[n]; fy += F[N+n]; fz += F[2ull*N+n]; } /* FORCE_FIELD, :1497-1503 / :1821-1827 */
	*fxp = fx; *fyp = fy; *fzp = fz;
	if(g->features&ORC_VOLUME_FORCE) { /* :1552-1555 / :1851-1854 */
		const float rho2 = 0.5f/(*rhon);
		*uxn = clamp_c(fmaf(fx, rho2, *uxn)); *uyn = clamp_c(fmaf(fy, rho2, *uyn)); *uzn = clamp_c(fmaf(fz, rho2, *uzn));
	} else { *uxn = clamp_c(*uxn); *uyn = clamp_c(*uyn); *uzn = clamp_c(*uzn); }
	return 1;
}

/* ---- stream_collide: src/kernel.cpp:1454-1636 (north_star subset) ---- */
void orc_stream_collide_F(const orc_grid* g, void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t, float fx0, float fy0, float fz0, const float* F) {
	const uint64_t N = cells_of(g);
	const uint32_t Q = g->Q;
	const int eb = (g->features&ORC_EQUILIBRIUM_BOUNDARIES)!=0u, vf = (g->features&ORC_VOLUME_FORCE)!=0u;
	ORC_PARALLEL_FOR
	for(uint64_t n=0ull; n<N; n++) {
		uint64_t j[QMAX]; float fhn[QMAX], feq[QMAX], Fin[QMAX];
		float rhon, uxn, uyn, uzn; uint8_t fb;
		float fx = fx0, fy = fy0, fz = fz0;
		if(!cell_front(g, fi, rho, u, flags, n, t, &fx, &fy, &fz, F, 1, j, fhn, &rhon, &uxn, &uyn, &uzn, &fb)) continue;
		if(vf) forcing_terms(g, uxn, uyn, uzn, fx, fy, fz, Fin); else for(uint32_t i=0u; i<Q; i++) Fin[i] = 0.0f;
		const int is_e = eb && fb==TYPE_E;
		if((g->features&ORC_UPDATE_FIELDS) && !is_e) { rho[n] = rhon; u[n] = uxn; u[N+n] = uyn; u[2ull*N+n] = uzn; } /* :1565-1573 */
		equilibrium(g, rhon, uxn, uyn, uzn, feq);
		float w = g->w;
		if(g->features&ORC_SUBGRID) { /* Smagorinsky-Lilly subgrid model, :1579-1593: relaxation rate from the non-equilibrium stress tensor */
			const float tau0 = 1.0f/w;
			float Hxx = 0.0f, Hyy = 0.0f, Hzz = 0.0f, Hxy = 0.0f, Hxz = 0.0f, Hyz = 0.0f;
			for(uint32_t i=1u; i<Q; i++) {
				const float fneqi = fhn[i]-feq[i];
				const float cxi = (float)EX[i], cyi = (float)EY[i], czi = (float)EZ[i];
				Hxx += cxi*cxi*fneqi;
				Hxy += cxi*cyi*fneqi; Hyy += cyi*cyi*fneqi;
				Hxz += cxi*czi*fneqi; Hyz += cyi*czi*fneqi; Hzz += czi*czi*fneqi;
			}
			const float Qs = Hxx*Hxx+Hyy*Hyy+Hzz*Hzz+2.0f*(Hxy*Hxy+Hxz*Hxz+Hyz*Hyz);
			w = 2.0f/(tau0+sqrtf(tau0*tau0+0.76421222f*sqrtf(Qs)/rhon));
		}
		if(g->collision==ORC_SRT) { /* :1595-1604 */
			if(vf) { const float c_tau = fmaf(w, -0.5f, 1.0f); for(uint32_t i=0u; i<Q; i++) Fin[i] *= c_tau; }
			for(uint32_t i=0u; i<Q; i++) fhn[i] = is_e ? feq[i] : fmaf(1.0f-w, fhn[i], fmaf(w, feq[i], Fin[i]));
		} else { /* TRT :1605-1633 */
			const float wp = w;
			const float wm = 1.0f/(0.1875f/(1.0f/w-0.5f)+0.5f);
			if(vf) {
				const float c_taup = fmaf(wp, -0.25f, 0.5f), c_taum = fmaf(wm, -0.25f, 0.5f);
				float Fib[QMAX];
				Fib[0] = Fin[0];
				for(uint32_t i=1u; i<Q; i+=2u) { Fib[i] = Fin[i+1u]; Fib[i+1u] = Fin[i]; }
				for(uint32_t i=0u; i<Q; i++) Fin[i] = fmaf(c_taup, Fin[i]+Fib[i], c_taum*(Fin[i]-Fib[i]));
			}
			float fhb[QMAX], feb[QMAX];
			fhb[0] = fhn[0]; feb[0] = feq[0];
			for(uint32_t i=1u; i<Q; i+=2u) { fhb[i] = fhn[i+1u]; fhb[i+1u] = fhn[i]; feb[i] = feq[i+1u]; feb[i+1u] = feq[i]; }
			for(uint32_t i=0u; i<Q; i++) fhn[i] = is_e ? feq[i] :
				fmaf(0.5f*wp, feq[i]-fhn[i]+feb[i]-fhb[i], fmaf(0.5f*wm, feq[i]-feb[i]-fhn[i]+fhb[i], fhn[i]+Fin[i]));
		}
		push_ddfs(g, n, fhn, fi, j, t);
	}
}

void orc_stream_collide(const orc_grid* g, void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t, float fx, float fy, float fz) {
	orc_stream_collide_F(g, fi, rho, u, flags, t, fx, fy, fz, NULL);
}

/* ---- update_fields: src/kernel.cpp:1794-1870 ---- */
void orc_update_fields_F(const orc_grid* g, const void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t, float fx0, float fy0, float fz0, const float* F) {
	const uint64_t N = cells_of(g);
	const int eb = (g->features&ORC_EQUILIBRIUM_BOUNDARIES)!=0u;
	ORC_PARALLEL_FOR
	for(uint64_t n=0ull; n<N; n++) {
		uint64_t j[QMAX]; float fhn[QMAX];
		float rhon, uxn, uyn, uzn; uint8_t fb;
		float fx = fx0, fy = fy0, fz = fz0;
		if(!cell_front(g, fi, rho, u, flags, n, t, &fx, &fy, &fz, F, 0, j, fhn, &rhon, &uxn, &uyn, &uzn, &fb)) continue;
		if(eb && fb==TYPE_E) continue; /* :1862-1864 */
		rho[n] = rhon; u[n] = uxn; u[N+n] = uyn; u[2ull*N+n] = uzn;
	}
}
void orc_update_fields(const orc_grid* g, const void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t, float fx, float fy, float fz) {
	orc_update_fields_F(g, fi, rho, u, flags, t, fx, fy, fz, NULL);
}

/* ================================================================================================================
 * FORCE_FIELD: src/kernel.cpp:1873-1959, host side src/lbm.cpp:206-239,986-1016
 * ================================================================================================================ */
/* update_force_field, :1873-1884: force of the fluid on every solid cell = twice the momentum of the populations that stream into it
   (they are bounced back), via calculate_rho_u -> F = 2*Fb*(fx,fy,fz) with Fb the "density" and f the "velocity" */
void orc_update_force_field(const orc_grid* g, const void* fi, const uint8_t* flags, uint64_t t, float* F) {
	const uint64_t N = cells_of(g);
	ORC_PARALLEL_FOR
	for(uint64_t n=0ull; n<N; n++) {
		const xyz_t c = coords_of(g, n);
		if(halo_cell(g, c)) continue;
		if((flags[n]&TYPE_BO)!=TYPE_S) continue;
		uint64_t j[QMAX]; float fhn[QMAX];
		neighbours_of(g, n, j);
		pull_ddfs(g, n, fhn, fi, j, t);
		float Fb, fx, fy, fz;
		moments(g, fhn, &Fb, &fx, &fy, &fz);
		const float s = 2.0f*Fb; /* 2.0f*Fb*(float3)(fx,fy,fz): scalar product first, then scalar times vector */
		F[n] = s*fx; F[N+n] = s*fy; F[2ull*N+n] = s*fz;
	}
}
void orc_reset_force_field(const orc_grid* g, float* F) { /* :1885-1889, halo included */
	const uint64_t N = cells_of(g);
	for(uint64_t n=0ull; n<3ull*N; n++) F[n] = 0.0f;
}
/* object_center_of_mass / object_force / object_torque, :1901-1959. The reference reduces 3 floats per work-group in local memory with a
   stride-doubling tree (cache[lid] += cache[lid+s] for lid%(2s)==0, s = 1,2,4,..) and then adds the group results into object_sum with
   floating-point ATOMICS, i.e. in no defined order. The oracle fixes that order: the same tree per group of `group` consecutive cells
   (cl_workgroup_size), then the non-zero group sums added in ascending group order. The reference's result is one of the orderings of the
   same group sums; tests compare against it with a tolerance and against this function bit for bit. kind: 0 centre of mass (out[3] = cell
   count as raw bits), 1 force, 2 torque about (cx,cy,cz). position() = coordinates + 0.5 - 0.5*N (:827-829), local to the domain. */
static void tree_reduce3(float* cache, uint32_t group) {
	for(uint32_t s=1u; s<group; s*=2u) for(uint32_t lid=0u; lid+s<group; lid+=2u*s) { cache[3u*lid] += cache[3u*(lid+s)]; cache[3u*lid+1u] += cache[3u*(lid+s)+1u]; cache[3u*lid+2u] += cache[3u*(lid+s)+2u]; }
}
void orc_object_sum(const orc_grid* g, uint32_t kind, const float* F, const uint8_t* flags, uint8_t flag_marker, float cx, float cy, float cz, uint32_t group, float* out4) {
	const uint64_t N = cells_of(g);
	float sum[3] = { 0.0f, 0.0f, 0.0f }; uint32_t count = 0u;
	float* cache = (float*)malloc(sizeof(float)*3u*group);
	for(uint64_t base=0ull; base<N; base+=group) { /* one work-group; the NDRange is padded to a multiple of the group size with idle items */
		uint32_t cells = 0u;
		for(uint32_t lid=0u; lid<group; lid++) {
			const uint64_t n = base+lid;
			float v[3] = { 0.0f, 0.0f, 0.0f };
			if(n<N && flags[n]==flag_marker) {
				const xyz_t c = coords_of(g, n);
				const float px = (float)c.x+0.5f-0.5f*(float)g->Nx, py = (float)c.y+0.5f-0.5f*(float)g->Ny, pz = (float)c.z+0.5f-0.5f*(float)g->Nz;
				if(kind==0u) { v[0] = px; v[1] = py; v[2] = pz; cells++; }
				else if(kind==1u) { v[0] = F[n]; v[1] = F[N+n]; v[2] = F[2ull*N+n]; }
				else { /* cross(position-centre, F), src/kernel.cpp:1947 */
					const float rx = px-cx, ry = py-cy, rz = pz-cz, Fx = F[n], Fy = F[N+n], Fz = F[2ull*N+n];
					v[0] = ry*Fz-rz*Fy; v[1] = rz*Fx-rx*Fz; v[2] = rx*Fy-ry*Fx;
				}
			}
			cache[3u*lid] = v[0]; cache[3u*lid+1u] = v[1]; cache[3u*lid+2u] = v[2];
		}
		tree_reduce3(cache, group);
		if(kind==0u) { if(cells>0u) { sum[0] += cache[0]; sum[1] += cache[1]; sum[2] += cache[2]; count += cells; } }
		else { if(cache[0]!=0.0f) sum[0] += cache[0]; if(cache[1]!=0.0f) sum[1] += cache[1]; if(cache[2]!=0.0f) sum[2] += cache[2]; }
	}
	free(cache);
	out4[0] = sum[0]; out4[1] = sum[1]; out4[2] = sum[2]; out4[3] = float_of(count);
}
/* ---- halo transfer: src/kernel.cpp:2049-2158, host side src/lbm.cpp:1308-1354 ---- */
uint32_t orc_transfers(const orc_grid* g) { return g->Q==19u ? 5u : 9u; } /* src/lbm.cpp:7-23 */
uint64_t orc_area(const orc_grid* g, uint32_t axis) {
	return axis==0u ? (uint64_t)g->Ny*g->Nz : axis==1u ? (uint64_t)g->Nz*g->Nx : (uint64_t)g->Nx*g->Ny;
}
/* face cell a on layer `layer` of `axis`: decomposition of a differs per axis (:2053-2068) */
static inline uint64_t face_cell(const orc_grid* g, uint32_t axis, uint64_t a, uint32_t layer) {
	switch(axis) {
		case 0u: return index_of(g, layer, (uint32_t)(a%g->Ny), (uint32_t)(a/g->Ny));
		case 1u: return index_of(g, (uint32_t)(a/g->Nz), layer, (uint32_t)(a%g->Nz));
		default: return index_of(g, (uint32_t)(a%g->Nx), (uint32_t)(a/g->Nx), layer);
	}
}
/* directions crossing a face, listed so that position b pairs opposite directions on the two sides (:2069-2101).
   For each axis the "+" side holds the directions with a positive component along it; within a side the
   order is the reference's. */
static const uint8_t XFER19[6][5] = {
	{ 1, 7,13, 9,15 }, { 2, 8,14,10,16 },
	{ 3, 7,14,11,17 }, { 4, 8,13,12,18 },
	{ 5, 9,16,11,18 }, { 6,10,15,12,17 } };
static const uint8_t XFER27[6][9] = {
	{ 1, 7,13, 9,15,19,26,21,23 }, { 2, 8,14,10,16,20,25,22,24 },
	{ 3, 7,14,11,17,19,24,21,25 }, { 4, 8,13,12,18,20,23,22,26 },
	{ 5, 9,16,11,18,19,22,23,25 }, { 6,10,15,12,17,20,21,24,26 } };
static inline uint32_t xfer_dir(const orc_grid* g, uint32_t side, uint32_t b) { return g->Q==19u ? XFER19[side][b] : XFER27[side][b]; }

static inline void copy_raw(const orc_grid* g, void* dst, uint64_t di, const void* src, uint64_t si) { /* fpxx_copy: raw bits */
	if(g->storage==ORC_FP32) ((uint32_t*)dst)[di] = ((const uint32_t*)src)[si];
	else ((uint16_t*)dst)[di] = ((const uint16_t*)src)[si];
}
static void extract_side(const orc_grid* g, uint64_t a, uint64_t A, uint64_t n, uint32_t side, uint64_t t, void* buf, const void* fi) { /* :2102-2110 */
	uint64_t j[QMAX]; neighbours_of(g, n, j);
	const uint32_t odd = (uint32_t)(t&1ull), T = orc_transfers(g);
	for(uint32_t b=0u; b<T; b++) {
		const uint32_t i = xfer_dir(g, side, b);
		const uint64_t cell = (i&1u) ? j[i] : n;
		const uint32_t slot = odd ? ((i&1u) ? i+1u : i-1u) : i;
		copy_raw(g, buf, (uint64_t)b*A+a, fi, slot_index(g, cell, slot));
	}
}
static void insert_side(const orc_grid* g, uint64_t a, uint64_t A, uint64_t n, uint32_t side, uint64_t t, const void* buf, void* fi) { /* :2111-2119 */
	uint64_t j[QMAX]; neighbours_of(g, n, j);
	const uint32_t odd = (uint32_t)(t&1ull), T = orc_transfers(g);
	for(uint32_t b=0u; b<T; b++) {
		const uint32_t i = xfer_dir(g, side, b);
		const uint64_t cell = (i&1u) ? n : j[i-1u];
		const uint32_t slot = odd ? i : ((i&1u) ? i+1u : i-1u);
		copy_raw(g, fi, slot_index(g, cell, slot), buf, (uint64_t)b*A+a);
	}
}
static inline uint32_t axis_len(const orc_grid* g, uint32_t axis) { return axis==0u ? g->Nx : axis==1u ? g->Ny : g->Nz; }

void orc_transfer_extract_fi(const orc_grid* g, uint32_t axis, uint64_t t, void* buf_p, void* buf_m, const void* fi) { /* :2120-2125 */
	const uint64_t A = orc_area(g, axis); const uint32_t L = axis_len(g, axis);
	for(uint64_t a=0ull; a<A; a++) {
		extract_side(g, a, A, face_cell(g, axis, a, L-2u), 2u*axis+0u, t, buf_p, fi);
		extract_side(g, a, A, face_cell(g, axis, a, 1u   ), 2u*axis+1u, t, buf_m, fi);
	}
}
void orc_transfer_insert_fi(const orc_grid* g, uint32_t axis, uint64_t t, const void* buf_p, const void* buf_m, void* fi) { /* :2126-2131 */
	const uint64_t A = orc_area(g, axis); const uint32_t L = axis_len(g, axis);
	for(uint64_t a=0ull; a<A; a++) {
		insert_side(g, a, A, face_cell(g, axis, a, L-1u), 2u*axis+0u, t, buf_p, fi);
		insert_side(g, a, A, face_cell(g, axis, a, 0u   ), 2u*axis+1u, t, buf_m, fi);
	}
}
/* rho/u/flags halo: 4 float planes then one byte plane at byte offset 16*A (:2133-2158) */
static void extract_ruf(uint64_t a, uint64_t A, uint64_t n, uint64_t N, void* buf, const float* rho, const float* u, const uint8_t* flags) {
	float* fb = (float*)buf;
	fb[a] = rho[n]; fb[A+a] = u[n]; fb[2ull*A+a] = u[N+n]; fb[3ull*A+a] = u[2ull*N+n];
	((uint8_t*)buf)[16ull*A+a] = flags[n];
}
static void insert_ruf(uint64_t a, uint64_t A, uint64_t n, uint64_t N, const void* buf, float* rho, float* u, uint8_t* flags) {
	const float* fb = (const float*)buf;
	rho[n] = fb[a]; u[n] = fb[A+a]; u[N+n] = fb[2ull*A+a]; u[2ull*N+n] = fb[3ull*A+a];
	flags[n] = ((const uint8_t*)buf)[16ull*A+a];
}
void orc_transfer_extract_rho_u_flags(const orc_grid* g, uint32_t axis, uint64_t t, void* buf_p, void* buf_m, const float* rho, const float* u, const uint8_t* flags) {
	const uint64_t A = orc_area(g, axis), N = cells_of(g); const uint32_t L = axis_len(g, axis);
	(void)t;
	for(uint64_t a=0ull; a<A; a++) {
		extract_ruf(a, A, face_cell(g, axis, a, L-2u), N, buf_p, rho, u, flags);
		extract_ruf(a, A, face_cell(g, axis, a, 1u   ), N, buf_m, rho, u, flags);
	}
}
void orc_transfer_insert_rho_u_flags(const orc_grid* g, uint32_t axis, uint64_t t, const void* buf_p, const void* buf_m, float* rho, float* u, uint8_t* flags) {
	const uint64_t A = orc_area(g, axis), N = cells_of(g); const uint32_t L = axis_len(g, axis);
	(void)t;
	for(uint64_t a=0ull; a<A; a++) {
		insert_ruf(a, A, face_cell(g, axis, a, L-1u), N, buf_p, rho, u, flags);
		insert_ruf(a, A, face_cell(g, axis, a, 0u   ), N, buf_m, rho, u, flags);
	}
}

/* transfer_extract_F / transfer__insert_F, :2173-2196: three float planes per side */
void orc_transfer_extract_F(const orc_grid* g, uint32_t axis, uint64_t t, void* buf_p, void* buf_m, const float* F) {
	const uint64_t A = orc_area(g, axis), N = cells_of(g); const uint32_t L = axis_len(g, axis);
	(void)t;
	for(uint64_t a=0ull; a<A; a++) {
		const uint64_t np = face_cell(g, axis, a, L-2u), nm = face_cell(g, axis, a, 1u);
		for(uint64_t k=0ull; k<3ull; k++) { ((float*)buf_p)[k*A+a] = F[k*N+np]; ((float*)buf_m)[k*A+a] = F[k*N+nm]; }
	}
}
void orc_transfer_insert_F(const orc_grid* g, uint32_t axis, uint64_t t, const void* buf_p, const void* buf_m, float* F) {
	const uint64_t A = orc_area(g, axis), N = cells_of(g); const uint32_t L = axis_len(g, axis);
	(void)t;
	for(uint64_t a=0ull; a<A; a++) {
		const uint64_t np = face_cell(g, axis, a, L-1u), nm = face_cell(g, axis, a, 0u);
		for(uint64_t k=0ull; k<3ull; k++) { F[k*N+np] = ((const float*)buf_p)[k*A+a]; F[k*N+nm] = ((const float*)buf_m)[k*A+a]; }
	}
}


/* ================================================================================================================
 * voxelize_mesh / unvoxelize_mesh -- src/kernel.cpp:2267-2357. One work item per cell of the face normal to
 * `direction` casts a ray along the axis through the column, intersects it with every triangle (bidirectional Moeller-Trumbore),
 * sorts the hit distances and walks the column toggling inside/outside. cross() and dot() are the plain expressions
 * (a.y*b.z-a.z*b.y, ...; a.x*b.x+a.y*b.y+a.z*b.z, left to right), every operation separately rounded.
 * Ox,Oy,Oz: offset of this domain in the global grid (def_Ox.., src/lbm.cpp:335); bbu: the 16 floats of
 * LBM_Domain::voxelize_mesh_on_device (src/lbm.cpp:279-296): triangle count (as bits), bounding box, rotation centre, linear and
 * rotational velocity.
 * ================================================================================================================ */
typedef struct { float x, y, z; } f3_t;
static inline f3_t f3(float x, float y, float z) { f3_t r = { x, y, z }; return r; }
static inline f3_t f3_sub(f3_t a, f3_t b) { return f3(a.x-b.x, a.y-b.y, a.z-b.z); }
static inline f3_t f3_add(f3_t a, f3_t b) { return f3(a.x+b.x, a.y+b.y, a.z+b.z); }
static inline f3_t f3_cross(f3_t a, f3_t b) { return f3(a.y*b.z-a.z*b.y, a.z*b.x-a.x*b.z, a.x*b.y-a.y*b.x); }
static inline float f3_dot(f3_t a, f3_t b) { return a.x*b.x+a.y*b.y+a.z*b.z; }
static inline int clampi(int v, int lo, int hi) { return v<lo ? lo : v>hi ? hi : v; }
static inline f3_t position_of(const orc_grid* g, uint32_t x, uint32_t y, uint32_t z) { /* src/kernel.cpp:828-830 */
	return f3((float)x+0.5f-0.5f*(float)g->Nx, (float)y+0.5f-0.5f*(float)g->Ny, (float)z+0.5f-0.5f*(float)g->Nz);
}
void orc_voxelize_mesh(const orc_grid* g, int Ox, int Oy, int Oz, uint32_t direction, void* fi, float* u, uint8_t* flags, uint64_t t, uint8_t flag,
	const float* p0, const float* p1, const float* p2, const float* bbu) {
	const uint64_t A = orc_area(g, direction), N = cells_of(g);
	const uint32_t triangle_number = bits_of(bbu[0]);
	const float x0 = bbu[1], y0 = bbu[2], z0 = bbu[3], x1 = bbu[4], y1 = bbu[5], z1 = bbu[6];
	const float cx = bbu[7], cy = bbu[8], cz = bbu[9], ux = bbu[10], uy = bbu[11], uz = bbu[12], rx = bbu[13], ry = bbu[14], rz = bbu[15];
	const f3_t offset = f3(0.5f*(float)((int)g->Nx+2*Ox)-0.5f, 0.5f*(float)((int)g->Ny+2*Oy)-0.5f, 0.5f*(float)((int)g->Nz+2*Oz)-0.5f);
	ORC_PARALLEL_FOR
	for(uint64_t a=0u; a<A; a++) {
		const uint32_t hmin = direction==0u ? (uint32_t)clampi((int)x0-Ox, 0, (int)g->Nx-1) : direction==1u ? (uint32_t)clampi((int)y0-Oy, 0, (int)g->Ny-1) : (uint32_t)clampi((int)z0-Oz, 0, (int)g->Nz-1);
		const uint32_t hmax = direction==0u ? (uint32_t)clampi((int)x1-Ox, 0, (int)g->Nx-1) : direction==1u ? (uint32_t)clampi((int)y1-Oy, 0, (int)g->Ny-1) : (uint32_t)clampi((int)z1-Oz, 0, (int)g->Nz-1);
		uint32_t X, Y, Z;
		if(direction==0u) { X = hmin; Y = (uint32_t)(a%g->Ny); Z = (uint32_t)(a/g->Ny); }
		else if(direction==1u) { X = (uint32_t)(a/g->Nz); Y = hmin; Z = (uint32_t)(a%g->Nz); }
		else { X = (uint32_t)(a%g->Nx); Y = (uint32_t)(a/g->Nx); Z = hmin; }
		const f3_t r_origin = f3_add(position_of(g, X, Y, Z), offset);
		const f3_t r_direction = f3((float)(direction==0u), (float)(direction==1u), (float)(direction==2u));
		uint32_t intersections = 0u, intersections_check = 0u;
		uint16_t distances[64];
		const int outside_box = direction==0u ? (r_origin.y<y0||r_origin.z<z0||r_origin.y>=y1||r_origin.z>=z1) : direction==1u ? (r_origin.x<x0||r_origin.z<z0||r_origin.x>=x1||r_origin.z>=z1) : (r_origin.x<x0||r_origin.y<y0||r_origin.x>=x1||r_origin.y>=y1);
		if(outside_box) continue;
		for(uint32_t i=0u; i<triangle_number; i++) {
			const f3_t p0i = f3(p0[3u*i], p0[3u*i+1u], p0[3u*i+2u]), p1i = f3(p1[3u*i], p1[3u*i+1u], p1[3u*i+2u]), p2i = f3(p2[3u*i], p2[3u*i+1u], p2[3u*i+2u]);
			const f3_t eu = f3_sub(p1i, p0i), ev = f3_sub(p2i, p0i), ew = f3_sub(r_origin, p0i), h = f3_cross(r_direction, ev), q = f3_cross(ew, eu);
			const float gdet = f3_dot(eu, h), f = 1.0f/gdet, s = f*f3_dot(ew, h), tt = f*f3_dot(r_direction, q), d = f*f3_dot(ev, q);
			if(gdet!=0.0f&&s>=0.0f&&s<1.0f&&tt>=0.0f&&s+tt<1.0f) {
				if(d>0.0f) {
					if(intersections<64u&&d<65536.0f) distances[intersections] = (uint16_t)d;
					intersections++;
				} else intersections_check++;
			}
		}
		for(uint32_t i=1u; i<(intersections<64u ? intersections : 64u); i++) { /* insertion sort */
			const uint16_t tv = distances[i];
			uint32_t j = i;
			while(j>0u&&distances[j-1u]>tv) { distances[j] = distances[j-1u]; j--; }
			distances[j] = tv;
		}
		int inside = (intersections%2u)&&(intersections_check%2u);
		const int set_u = ux*ux+uy*uy+uz*uz+rx*rx+ry*ry+rz*rz>0.0f;
		uint32_t intersection = intersections%2u!=intersections_check%2u;
		const uint32_t h0 = direction==0u ? X : direction==1u ? Y : Z;
		const uint32_t last = intersections-1u<63u ? intersections-1u : 63u; /* min(intersections-1u, 63u) with unsigned wrap for 0 */
		const uint32_t hmesh = h0+(uint32_t)(intersections>0u ? distances[last] : distances[63]); /* (value unused when there are no intersections: inside stays 0) */
		for(uint32_t hh=h0; hh<=hmax; hh++) {
			while(intersection<intersections&&hh>h0+(uint32_t)distances[intersection<63u ? intersection : 63u]) { inside = !inside; intersection++; }
			inside = inside&&(intersection<intersections&&hh<hmesh);
			const uint64_t n = index_of(g, direction==0u ? hh : X, direction==1u ? hh : Y, direction==2u ? hh : Z);
			uint8_t flagsn = flags[n];
			const xyz_t c = coords_of(g, n);
			const f3_t p = f3_add(position_of(g, c.x, c.y, c.z), offset);
			const f3_t u_set = f3_add(f3(ux, uy, uz), f3_cross(f3_sub(f3(cx, cy, cz), p), f3(rx, ry, rz)));
			if(inside) {
				flagsn = (uint8_t)((flagsn&~TYPE_BO)|flag);
				if(set_u) { u[n] = u_set.x; u[N+n] = u_set.y; u[2u*N+n] = u_set.z; }
			} else if((flagsn&TYPE_BO)==TYPE_S&&(flagsn&0xC0u)==(flag&0xC0u)) { /* TYPE_XY = TYPE_X|TYPE_Y */
				const float unx = u[n], uny = u[N+n], unz = u[2u*N+n];
				if(unx==u_set.x&&uny==u_set.y&&unz==u_set.z) {
					if(set_u) { /* the solid cell becomes fluid again: its DDFs restart from equilibrium at rho=1 */
						uint64_t j[QMAX]; float feq[QMAX];
						neighbours_of(g, n, j);
						equilibrium(g, 1.0f, unx, uny, unz, feq);
						push_ddfs(g, n, feq, fi, j, t);
					}
					flagsn = (flagsn&TYPE_BO)==TYPE_MS ? (uint8_t)(flagsn&~TYPE_MS) : (uint8_t)(flagsn&~flag);
				}
			}
			flags[n] = flagsn;
		}
	}
}
void orc_unvoxelize_mesh(const orc_grid* g, int Ox, int Oy, int Oz, uint8_t* flags, uint8_t flag, float x0, float y0, float z0, float x1, float y1, float z1) { /* :2351-2357 */
	const uint64_t N = cells_of(g);
	const f3_t offset = f3(0.5f*(float)((int)g->Nx+2*Ox)-0.5f, 0.5f*(float)((int)g->Ny+2*Oy)-0.5f, 0.5f*(float)((int)g->Nz+2*Oz)-0.5f);
	ORC_PARALLEL_FOR
	for(uint64_t n=0u; n<N; n++) {
		const xyz_t c = coords_of(g, n);
		const f3_t p = f3_add(position_of(g, c.x, c.y, c.z), offset);
		if(p.x>=x0-1.0f&&p.y>=y0-1.0f&&p.z>=z0-1.0f&&p.x<=x1+1.0f&&p.y<=y1+1.0f&&p.z<=z1+1.0f) flags[n] &= (uint8_t)~flag;
	}
}
