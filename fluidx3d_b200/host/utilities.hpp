// utilities.hpp -- the small part of the reference's general utility header that the LBM host surface and the scene
// scripts on the hot path rely on (FluidX3D v3.7 src/utilities.hpp: typedefs, parallel_for :60-93, Clock :95-103,
// float3/uint3 helpers, console messages :4042-4092). Written fresh; only names and call signatures are kept so that
// setup.cpp-style scenes compile unchanged. Image IO, mesh IO, colour helpers etc. are outside the path.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <string>
#include <thread>
#include <vector>
using std::string;
using std::vector;
using std::thread;
using std::to_string;
using std::min;
using std::max;

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef int64_t slong;
typedef uint64_t ulong;

#define pif 3.1415927f
#define pi 3.141592653589793
#define max_uint 4294967295u
#define max_ulong 18446744073709551615ull
#define max_float 3.402823466E38f
#define max_double 1.7976931348623158E308

inline float as_float(const uint x) { float f; std::memcpy(&f, &x, 4); return f; }
inline uint as_uint(const float x) { uint u; std::memcpy(&u, &x, 4); return u; }
inline float sq(const float x) { return x*x; }
inline double sq(const double x) { return x*x; }
inline uint sq(const uint x) { return x*x; }
inline int sq(const int x) { return x*x; }
inline float cb(const float x) { return x*x*x; }
inline double cb(const double x) { return x*x*x; }
inline uint to_uint(const float x) { return (uint)fmax(x+0.5f, 0.5f); }
inline uint to_uint(const double x) { return (uint)fmax(x+0.5, 0.5); }
inline uint gcd(uint x, uint y) { while(y!=0u) { const uint t = x%y; x = y; y = t; } return x; }
inline uint lcm(const uint x, const uint y) { return x/gcd(x, y)*y; }

// ---- run a loop body over [0,N) on all hardware threads; same call forms as the reference ----
inline void parallel_for(const ulong N, const uint threads, std::function<void(ulong, uint)> body) {
	const uint T = threads>0u ? threads : 1u;
	vector<thread> pool;
	pool.reserve(T);
	for(uint t=0u; t<T; t++) pool.emplace_back([=]() { for(ulong n=N*(ulong)t/(ulong)T; n<N*(ulong)(t+1u)/(ulong)T; n++) body(n, t); });
	for(thread& w : pool) w.join();
}
inline void parallel_for(const ulong N, const uint threads, std::function<void(ulong)> body) { parallel_for(N, threads, [&](ulong n, uint) { body(n); }); }
inline void parallel_for(const ulong N, std::function<void(ulong)> body) { parallel_for(N, max(1u, (uint)thread::hardware_concurrency()), body); }

class Clock {
	std::chrono::steady_clock::time_point t0;
public:
	Clock() { start(); }
	void start() { t0 = std::chrono::steady_clock::now(); }
	double stop() const { return std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count(); }
};
inline void sleep(const double seconds) { if(seconds>0.0) std::this_thread::sleep_for(std::chrono::duration<double>(seconds)); }

// ---- small vector types used by scenes: lbm.center(), lbm.size(), shape predicates ----
struct float3 {
	float x=0.0f, y=0.0f, z=0.0f;
	float3() {}
	float3(const float v) : x(v), y(v), z(v) {}
	float3(const float x_, const float y_, const float z_) : x(x_), y(y_), z(z_) {}
	template<class A, class B, class C> float3(const A x_, const B y_, const C z_) : x((float)x_), y((float)y_), z((float)z_) {}
	float3 operator+(const float3& o) const { return float3(x+o.x, y+o.y, z+o.z); }
	float3 operator-(const float3& o) const { return float3(x-o.x, y-o.y, z-o.z); }
	float3 operator*(const float s) const { return float3(x*s, y*s, z*s); }
	float3 operator/(const float s) const { return float3(x/s, y/s, z/s); }
	float3& operator+=(const float3& o) { x += o.x; y += o.y; z += o.z; return *this; }
	float3 operator-() const { return float3(-x, -y, -z); }
};
inline float3 operator*(const float s, const float3& v) { return v*s; }
inline float dot(const float3& a, const float3& b) { return a.x*b.x+a.y*b.y+a.z*b.z; }
inline float3 cross(const float3& a, const float3& b) { return float3(a.y*b.z-a.z*b.y, a.z*b.x-a.x*b.z, a.x*b.y-a.y*b.x); }
inline float length(const float3& v) { return sqrtf(dot(v, v)); }
inline float3 normalize(const float3& v) { const float l = length(v); return l>0.0f ? v/l : v; }
struct uint3 {
	uint x=0u, y=0u, z=0u;
	uint3() {}
	uint3(const uint v) : x(v), y(v), z(v) {}
	uint3(const uint x_, const uint y_, const uint z_) : x(x_), y(y_), z(z_) {}
	uint3 operator/(const uint s) const { return uint3(x/s, y/s, z/s); }
	uint3 operator*(const uint s) const { return uint3(x*s, y*s, z*s); }
};

// ---- number formatting: like the reference, floats print with 1+8 significant digits ----
inline string to_string(const float x, const uint decimals) { char b[64]; std::snprintf(b, sizeof(b), "%.*f", (int)decimals, (double)x); return b; }
inline string to_string(const double x, const uint decimals) { char b[64]; std::snprintf(b, sizeof(b), "%.*f", (int)decimals, x); return b; }
inline string alignl(const uint n, const string& s) { return s.size()<n ? s+string(n-s.size(), ' ') : s; }
inline string alignr(const uint n, const string& s) { return s.size()<n ? string(n-s.size(), ' ')+s : s; }
inline void print(const string& s) { std::cout << s; }
inline void println(const string& s="") { std::cout << s << std::endl; }

// ---- console messages; an error is fatal, as in the reference (print_error -> exit(1), src/utilities.hpp:4074-4087) ----
inline void print_message(const string& kind, const string& message) {
	const size_t width = 77u-kind.size()-2u; // boxed to the 79 column table the console output uses
	size_t p = 0u;
	bool first = true;
	while(p<message.size() || first) {
		const string part = message.substr(p, width);
		std::cout << "| " << (first ? kind+": " : string(kind.size()+2u, ' ')) << part << string(width-part.size(), ' ') << "|" << std::endl;
		p += width; first = false;
	}
}
inline void print_info(const string& s) { print_message("Info", s); }
inline void print_warning(const string& s) { print_message("Warning", s); }
[[noreturn]] inline void print_error(const string& s) {
	print_message("Error", s);
	std::cout << "|-----------------------------------------------------------------------------|" << std::endl;
	std::exit(1);
}
inline void wait() { std::cin.get(); }
inline string get_exe_path() { return "./"; }
