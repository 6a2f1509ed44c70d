// mesh.hpp -- triangle meshes for the GPU voxeliser: float3x3, Mesh, read_stl. Host side of LBM::voxelize_stl /
// voxelize_mesh_on_device (reference: FluidX3D v3.7 src/utilities.hpp:1048-1213 float3x3, :4440-4528 Mesh, :4530-4581 read_stl;
// src/lbm.cpp:275-327, 1074-1145). Same names and call signatures, so that scenes compile unchanged; binary .stl only, as there.
// Every coordinate is computed in binary32 in the reference's order (rotate, then centre + scale*(offset + p)): the voxelised
// flags must come out identical, and the ray/triangle tests downstream are sensitive to the last bit of a vertex.
#pragma once
#include "utilities.hpp"
#include <fstream>

struct float3x3 { // row-major 3x3 matrix; float3x3(1.0f) is the identity
	float xx=0.0f, xy=0.0f, xz=0.0f, yx=0.0f, yy=0.0f, yz=0.0f, zx=0.0f, zy=0.0f, zz=0.0f;
	float3x3() {}
	float3x3(const float diagonal) : xx(diagonal), yy(diagonal), zz(diagonal) {}
	float3x3(const float xx_, const float xy_, const float xz_, const float yx_, const float yy_, const float yz_, const float zx_, const float zy_, const float zz_)
		: xx(xx_), xy(xy_), xz(xz_), yx(yx_), yy(yy_), yz(yz_), zx(zx_), zy(zy_), zz(zz_) {}
	float3x3(const float3& axis, const float angle) { // rotation by `angle` radians about the normalised `axis` (Rodrigues)
		const float s = sinf(angle), c = cosf(angle), k = 1.0f-c;
		xx = sq(axis.x)+(1.0f-sq(axis.x))*c; xy = axis.x*axis.y*k-axis.z*s;     xz = axis.x*axis.z*k+axis.y*s;
		yx = axis.x*axis.y*k+axis.z*s;     yy = sq(axis.y)+(1.0f-sq(axis.y))*c; yz = axis.y*axis.z*k-axis.x*s;
		zx = axis.x*axis.z*k-axis.y*s;     zy = axis.y*axis.z*k+axis.x*s;     zz = sq(axis.z)+(1.0f-sq(axis.z))*c;
	}
};
inline float3 operator*(const float3x3& m, const float3& v) { return float3(m.xx*v.x+m.xy*v.y+m.xz*v.z, m.yx*v.x+m.yy*v.y+m.yz*v.z, m.zx*v.x+m.zy*v.y+m.zz*v.z); }
inline float3x3 operator*(const float3x3& a, const float3x3& b) {
	return float3x3(a.xx*b.xx+a.xy*b.yx+a.xz*b.zx, a.xx*b.xy+a.xy*b.yy+a.xz*b.zy, a.xx*b.xz+a.xy*b.yz+a.xz*b.zz,
	                a.yx*b.xx+a.yy*b.yx+a.yz*b.zx, a.yx*b.xy+a.yy*b.yy+a.yz*b.zy, a.yx*b.xz+a.yy*b.yz+a.yz*b.zz,
	                a.zx*b.xx+a.zy*b.yx+a.zz*b.zx, a.zx*b.xy+a.zy*b.yy+a.zz*b.zy, a.zx*b.xz+a.zy*b.yz+a.zz*b.zz);
}
inline float radians(const float degrees) { return (pif/180.0f)*degrees; }

struct Mesh { // closed triangle surface: vertex i of triangle k is p<i>[k]
	uint triangle_number = 0u;
	float3 center, pmin, pmax; // reference point for scale()/rotate(), bounding box
	vector<float3> p0, p1, p2;
	Mesh(const uint triangles, const float3& center_) : triangle_number(triangles), center(center_), pmin(center_), pmax(center_), p0(triangles), p1(triangles), p2(triangles) {}
	void find_bounds() {
		if(triangle_number==0u) return;
		pmin = pmax = p0[0];
		for(uint k=0u; k<triangle_number; k++) for(const float3* p : { &p0[k], &p1[k], &p2[k] }) {
			pmin = float3(fminf(pmin.x, p->x), fminf(pmin.y, p->y), fminf(pmin.z, p->z));
			pmax = float3(fmaxf(pmax.x, p->x), fmaxf(pmax.y, p->y), fmaxf(pmax.z, p->z));
		}
	}
	void scale(const float factor) { // about the centre
		for(vector<float3>* p : { &p0, &p1, &p2 }) for(float3& v : *p) v = factor*(v-center)+center;
		pmin = factor*(pmin-center)+center; pmax = factor*(pmax-center)+center;
	}
	void translate(const float3& translation) {
		for(vector<float3>* p : { &p0, &p1, &p2 }) for(float3& v : *p) v += translation;
		center += translation; pmin += translation; pmax += translation;
	}
	void rotate(const float3x3& rotation) { // about the centre
		for(vector<float3>* p : { &p0, &p1, &p2 }) for(float3& v : *p) v = rotation*(v-center)+center;
		find_bounds();
	}
	void set_center(const float3& c) { center = c; }
	const float3& get_center() const { return center; }
	float3 get_center_of_mass() const { // of the enclosed volume: signed tetrahedra against the origin
		double V = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
		for(uint k=0u; k<triangle_number; k++) {
			const double dV = (double)dot(p0[k], cross(p1[k], p2[k]))/6.0;
			const float3 avg = 0.25f*(p0[k]+p1[k]+p2[k]);
			V += dV; cx += dV*(double)avg.x; cy += dV*(double)avg.y; cz += dV*(double)avg.z;
		}
		return float3((float)(cx/V), (float)(cy/V), (float)(cz/V));
	}
	float3 get_bounding_box_size() const { return pmax-pmin; }
	float3 get_bounding_box_center() const { return 0.5f*(pmin+pmax); }
	float get_min_size() const { return fminf(fminf(pmax.x-pmin.x, pmax.y-pmin.y), pmax.z-pmin.z); }
	float get_max_size() const { return fmaxf(fmaxf(pmax.x-pmin.x, pmax.y-pmin.y), pmax.z-pmin.z); }
	float get_scale_for_box_fit(const float3& box_size) const { return fminf(fminf(box_size.x/(pmax.x-pmin.x), box_size.y/(pmax.y-pmin.y)), box_size.z/(pmax.z-pmin.z)); }
};

// binary .stl: 80-byte header, uint32 triangle count, then 50 bytes per triangle (normal, 3 vertices, 2 attribute bytes)
inline Mesh* read_stl_raw(const string& path, const bool reposition, const float3& box_size, const float3& center, const float3x3& rotation, const float size) {
	const string filename = path.size()>=4u && path.substr(path.size()-4u)==".stl" ? path : path+".stl";
	std::ifstream file(filename, std::ios::in|std::ios::binary);
	if(file.fail()) print_error("File \""+filename+"\" does not exist!");
	const vector<char> data((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
	if(data.size()<84u) print_error("File \""+filename+"\" is corrupt!");
	uint triangles = 0u;
	std::memcpy(&triangles, data.data()+80, 4);
	if(triangles==0u || data.size()!=84ull+50ull*(ulong)triangles) print_error("File \""+filename+"\" is corrupt or unsupported! Only binary .stl files are supported.");
	print_info("Loading \""+filename+"\" with "+to_string(triangles)+" triangles.");
	Mesh* mesh = new Mesh(triangles, center);
	for(uint k=0u; k<triangles; k++) {
		float v[12];
		std::memcpy(v, data.data()+84ull+50ull*(ulong)k, 48);
		mesh->p0[k] = rotation*float3(v[3], v[4], v[5]); mesh->p1[k] = rotation*float3(v[6], v[7], v[8]); mesh->p2[k] = rotation*float3(v[9], v[10], v[11]);
	}
	mesh->find_bounds();
	const float scale = size==0.0f ? mesh->get_scale_for_box_fit(box_size) : size>0.0f ? size/mesh->get_max_size() : -size; // fit / longest side = size / factor -size
	const float3 offset = reposition ? -0.5f*(mesh->pmin+mesh->pmax) : float3(0.0f); // bounding-box centre -> `center`
	for(vector<float3>* p : { &mesh->p0, &mesh->p1, &mesh->p2 }) for(float3& q : *p) q = center+scale*(offset+q);
	mesh->find_bounds();
	return mesh;
}
inline Mesh* read_stl(const string& path, const float3& box_size, const float3& center, const float3x3& rotation, const float size) { return read_stl_raw(path, true, box_size, center, rotation, size); }
inline Mesh* read_stl(const string& path, const float3& box_size, const float3& center, const float size) { return read_stl_raw(path, true, box_size, center, float3x3(1.0f), size); }
inline Mesh* read_stl(const string& path, const float scale=1.0f, const float3x3& rotation=float3x3(1.0f), const float3& offset=float3(0.0f)) { return read_stl_raw(path, false, float3(1.0f), offset, rotation, -fabsf(scale)); }
