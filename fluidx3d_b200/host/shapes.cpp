#include "shapes.hpp"
// every predicate answers "is lattice point (x,y,z) inside the body", p is the body centre in lattice coordinates
static inline float3 rel(const uint x, const uint y, const uint z, const float3& p) { return float3((float)x, (float)y, (float)z)-p; }
bool sphere(const uint x, const uint y, const uint z, const float3& p, const float r) { const float3 d = rel(x, y, z, p); return dot(d, d)<=r*r; }
bool ellipsoid(const uint x, const uint y, const uint z, const float3& p, const float3& r) { const float3 d = rel(x, y, z, p); return sq(d.x/r.x)+sq(d.y/r.y)+sq(d.z/r.z)<=1.0f; }
bool cuboid(const uint x, const uint y, const uint z, const float3& p, const float3& l) { const float3 d = rel(x, y, z, p); return fabsf(d.x)<=0.5f*l.x && fabsf(d.y)<=0.5f*l.y && fabsf(d.z)<=0.5f*l.z; }
bool cube(const uint x, const uint y, const uint z, const float3& p, const float l) { return cuboid(x, y, z, p, float3(l)); }
bool cylinder(const uint x, const uint y, const uint z, const float3& p, const float3& n, const float r) { // axis n through p, length |n|, radius r
	const float3 d = rel(x, y, z, p);
	const float along = dot(normalize(n), d);
	return dot(d, d)-along*along<=r*r && along*along<=sq(0.5f*length(n));
}
bool plane(const uint x, const uint y, const uint z, const float3& p, const float3& n) { return dot(rel(x, y, z, p), n)<=0.0f; }
