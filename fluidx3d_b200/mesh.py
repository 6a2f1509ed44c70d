"""Triangle meshes for the GPU voxeliser: the host side of LBM::voxelize_stl / voxelize_mesh_on_device (FluidX3D v3.7
src/utilities.hpp:4425-4581 Mesh + read_stl, src/lbm.cpp:275-327, 1074-1145). Binary STL only, like the reference. All geometry
arithmetic is binary32 in the reference's order, because the voxelised flags have to come out identical."""
import numpy as np

_f32 = np.float32


class Mesh:
    """struct Mesh (src/utilities.hpp:4425-4528): triangle vertices p0, p1, p2 (float32 [triangles, 3]), bounding box, centre"""

    def __init__(self, p0, p1, p2, center):
        self.p0, self.p1, self.p2 = (np.ascontiguousarray(a, dtype=_f32) for a in (p0, p1, p2))
        self.center = np.array(center, dtype=_f32)
        self.find_bounds()

    @property
    def triangle_number(self):
        return int(self.p0.shape[0])

    def find_bounds(self):
        self.pmin = np.minimum(np.minimum(self.p0.min(axis=0), self.p1.min(axis=0)), self.p2.min(axis=0))
        self.pmax = np.maximum(np.maximum(self.p0.max(axis=0), self.p1.max(axis=0)), self.p2.max(axis=0))

    def get_center(self): return self.center
    def get_bounding_box_size(self): return self.pmax - self.pmin
    def get_bounding_box_center(self): return _f32(0.5) * (self.pmin + self.pmax)
    def get_min_size(self): return float(np.min(self.pmax - self.pmin))
    def get_max_size(self): return float(np.max(self.pmax - self.pmin))
    def get_scale_for_box_fit(self, box_size): return float(np.min(np.asarray(box_size, _f32) / (self.pmax - self.pmin)))

    def scale(self, factor):  # about the centre, src/utilities.hpp:4465-4473
        f = _f32(factor)
        for p in (self.p0, self.p1, self.p2): p[...] = f * (p - self.center) + self.center
        self.pmin = f * (self.pmin - self.center) + self.center; self.pmax = f * (self.pmax - self.center) + self.center

    def translate(self, translation):  # src/utilities.hpp:4474-4483
        t = np.asarray(translation, _f32)
        for p in (self.p0, self.p1, self.p2): p += t
        self.center = self.center + t; self.pmin = self.pmin + t; self.pmax = self.pmax + t

    def rotate(self, rotation):  # about the centre, src/utilities.hpp:4484-4491
        for p in (self.p0, self.p1, self.p2): p[...] = _rotate(rotation, p - self.center) + self.center
        self.find_bounds()


def _rotate(rotation, p):
    """float3x3 * float3 row by row (src/utilities.hpp:1211): m.xx*v.x+m.xy*v.y+m.xz*v.z, evaluated left to right in binary32"""
    R = np.asarray(rotation, dtype=_f32)
    return np.stack([((R[r, 0] * p[:, 0]).astype(_f32) + (R[r, 1] * p[:, 1]).astype(_f32) + (R[r, 2] * p[:, 2]).astype(_f32)).astype(_f32) for r in range(3)], axis=1)


def read_stl(path, box_size, center, size, rotation=None, reposition=True):
    """read_stl_raw (src/utilities.hpp:4530-4571). size == 0: fit the bounding box into box_size; size > 0: longest bounding-box
    side becomes `size` cells; size < 0: scale by -size. reposition: centre the bounding box on `center`."""
    if not path.endswith(".stl"): path += ".stl"
    with open(path, "rb") as f: raw = f.read()
    if len(raw) < 84: raise ValueError(f'File "{path}" is corrupt!')
    n = int(np.frombuffer(raw, np.uint32, 1, 80)[0])
    if n == 0 or len(raw) != 84 + 50 * n: raise ValueError(f'File "{path}" is corrupt or unsupported! Only binary .stl files are supported.')
    rec = np.frombuffer(raw, np.uint8, 50 * n, 84).reshape(n, 50)[:, :48].copy().view(_f32).reshape(n, 12)
    vertices = [rec[:, 3:6].copy(), rec[:, 6:9].copy(), rec[:, 9:12].copy()]
    if rotation is not None: vertices = [_rotate(rotation, v) for v in vertices]
    mesh = Mesh(*vertices, center)
    if size == 0.0: scale = _f32(mesh.get_scale_for_box_fit(box_size))
    elif size > 0.0: scale = _f32(_f32(size) / _f32(mesh.get_max_size()))
    else: scale = _f32(-size)
    offset = (_f32(-0.5) * (mesh.pmin + mesh.pmax)).astype(_f32) if reposition else np.zeros(3, _f32)
    c = np.asarray(center, _f32)
    for p in (mesh.p0, mesh.p1, mesh.p2): p[...] = c + scale * (offset + p)
    mesh.find_bounds()
    return mesh


def voxelize_parameters(mesh, rotation_center, linear_velocity, rotational_velocity):
    """bounding_box_and_velocity[16] and the ray direction of LBM_Domain::voxelize_mesh_on_device (src/lbm.cpp:279-321)"""
    bbu = np.zeros(16, _f32)
    bbu[0] = np.array([mesh.triangle_number], np.uint32).view(_f32)[0]
    bbu[1:4] = mesh.pmin - _f32(2.0)  # tolerance of 2 cells for re-voxelisation of moving objects
    bbu[4:7] = mesh.pmax + _f32(2.0)
    bbu[7:10] = np.asarray(rotation_center, _f32); bbu[10:13] = np.asarray(linear_velocity, _f32); bbu[13:16] = np.asarray(rotational_velocity, _f32)
    x0, y0, z0, x1, y1, z1 = (bbu[k] for k in range(1, 7))
    rot = bbu[13:16]
    direction = 0
    if rot[0] == 0 and rot[1] == 0 and rot[2] == 0:  # the face of the bounding box with the smallest area: fewest rays
        area = [(y1 - y0) * (z1 - z0), (z1 - z0) * (x1 - x0), (x1 - x0) * (y1 - y0)]
        for i in (1, 2):
            if area[i] < area[direction]: direction = i
    else:  # along the rotation axis
        for i in (1, 2):
            if abs(rot[i]) > abs(rot[direction]): direction = i
    return bbu, direction
