// shapes.hpp -- host-side geometry predicates used when flagging cells in a scene (the reference's src/shapes.hpp)
#pragma once
#include "utilities.hpp"
bool sphere(const uint x, const uint y, const uint z, const float3& p, const float r);
bool ellipsoid(const uint x, const uint y, const uint z, const float3& p, const float3& r);
bool cube(const uint x, const uint y, const uint z, const float3& p, const float l);
bool cuboid(const uint x, const uint y, const uint z, const float3& p, const float3& l);
bool cylinder(const uint x, const uint y, const uint z, const float3& p, const float3& n, const float r);
bool plane(const uint x, const uint y, const uint z, const float3& p, const float3& n);
