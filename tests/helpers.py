"""Test helpers: ctypes bindings for the CPU oracle (oracle/_build) and for oracle/_ref (the reference's own
device code compiled natively), plus a backend-agnostic restatement of the reference's host sequencing
(LBM::initialize / do_time_step / communicate_field, src/lbm.cpp:881-953,1343-1390) so that oracle, _ref and
the CUDA library can all be driven through identical steps. Test infrastructure only."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liblbm_oracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

FP32, FP16S, FP16C = 0, 1, 2
SRT, TRT = 0, 1
VOLUME_FORCE, EQUILIBRIUM_BOUNDARIES, UPDATE_FIELDS, SUBGRID, MOVING_BOUNDARIES, FORCE_FIELD = 1, 2, 4, 8, 16, 32
TYPE_MS = 3
TYPE_S, TYPE_E = 1, 2
STORAGE_NAMES = {FP32: "fp32", FP16S: "fp16s", FP16C: "fp16c"}
COLL_NAMES = {SRT: "srt", TRT: "trt"}


def ddf_dtype(storage):
    return np.float32 if storage == FP32 else np.uint16


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ROOT, "oracle", "lbm_oracle.c")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return ORACLE_SO


class OrcGrid(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("Nx", "Ny", "Nz", "Dx", "Dy", "Dz", "Q", "collision", "storage", "features")] + [("w", C.c_float)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleBackend:
    """the plain-C restatement (oracle/lbm_oracle.c)"""
    name = "oracle"

    def __init__(self, Q=19, collision=SRT, storage=FP32, features=0, threads=0):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.orc_fp16s_encode.restype = C.c_uint16; L.orc_fp16s_encode.argtypes = [C.c_float]
        L.orc_fp16c_encode.restype = C.c_uint16; L.orc_fp16c_encode.argtypes = [C.c_float]
        L.orc_fp16s_decode.restype = C.c_float; L.orc_fp16s_decode.argtypes = [C.c_uint16]
        L.orc_fp16c_decode.restype = C.c_float; L.orc_fp16c_decode.argtypes = [C.c_uint16]
        L.orc_w_from_nu.restype = C.c_float; L.orc_w_from_nu.argtypes = [C.c_float]
        L.orc_float_to_string.argtypes = [C.c_float, C.c_char_p, C.c_int]
        L.orc_area.restype = C.c_uint64
        L.orc_set_threads(threads)
        self.Q, self.collision, self.storage, self.features = Q, collision, storage, features
        self.g = OrcGrid(1, 1, 1, 1, 1, 1, Q, collision, storage, features, 1.0)

    def set_grid(self, Nx, Ny, Nz, Dx, Dy, Dz, w):
        self.g = OrcGrid(Nx, Ny, Nz, Dx, Dy, Dz, self.Q, self.collision, self.storage, self.features, w)

    def w_from_nu(self, nu):
        return float(self.lib.orc_w_from_nu(C.c_float(nu)))

    def float_to_string(self, x):
        buf = C.create_string_buffer(64)
        self.lib.orc_float_to_string(C.c_float(x), buf, 64)
        return buf.value.decode()

    def initialize(self, fi, rho, u, flags):
        self.lib.orc_initialize(C.byref(self.g), _p(fi), _p(rho), _p(u), _p(flags))

    def stream_collide(self, fi, rho, u, flags, t, fx=0.0, fy=0.0, fz=0.0, F=None):
        if F is not None: self.lib.orc_stream_collide_F(C.byref(self.g), _p(fi), _p(rho), _p(u), _p(flags), C.c_uint64(t), C.c_float(fx), C.c_float(fy), C.c_float(fz), _p(F))
        else: self.lib.orc_stream_collide(C.byref(self.g), _p(fi), _p(rho), _p(u), _p(flags), C.c_uint64(t), C.c_float(fx), C.c_float(fy), C.c_float(fz))

    def update_fields(self, fi, rho, u, flags, t, fx=0.0, fy=0.0, fz=0.0, F=None):
        if F is not None: self.lib.orc_update_fields_F(C.byref(self.g), _p(fi), _p(rho), _p(u), _p(flags), C.c_uint64(t), C.c_float(fx), C.c_float(fy), C.c_float(fz), _p(F))
        else: self.lib.orc_update_fields(C.byref(self.g), _p(fi), _p(rho), _p(u), _p(flags), C.c_uint64(t), C.c_float(fx), C.c_float(fy), C.c_float(fz))

    # FORCE_FIELD
    def update_force_field(self, fi, flags, t, F):
        self.lib.orc_update_force_field(C.byref(self.g), _p(fi), _p(flags), C.c_uint64(t), _p(F))

    def reset_force_field(self, F):
        self.lib.orc_reset_force_field(C.byref(self.g), _p(F))

    def object_sum(self, kind, F, flags, marker, center=(0.0, 0.0, 0.0), group=64):
        """kind 0 centre of mass, 1 force, 2 torque; returns float32[4] (x, y, z, cell count as raw bits)"""
        out = np.zeros(4, np.float32)
        self.lib.orc_object_sum(C.byref(self.g), C.c_uint32(kind), _p(F) if F is not None else None, _p(flags), C.c_uint8(marker), C.c_float(center[0]), C.c_float(center[1]), C.c_float(center[2]), C.c_uint32(group), _p(out))
        return out

    def extract_F(self, axis, t, bp, bm, F):
        self.lib.orc_transfer_extract_F(C.byref(self.g), C.c_uint32(axis), C.c_uint64(t), _p(bp), _p(bm), _p(F))

    def insert_F(self, axis, t, bp, bm, F):
        self.lib.orc_transfer_insert_F(C.byref(self.g), C.c_uint32(axis), C.c_uint64(t), _p(bp), _p(bm), _p(F))

    def update_moving_boundaries(self, u, flags):
        self.lib.orc_update_moving_boundaries(C.byref(self.g), _p(u), _p(flags))

    def voxelize_mesh(self, offsets, direction, fi, u, flags, t, flag, p0, p1, p2, bbu):
        self.lib.orc_voxelize_mesh(C.byref(self.g), C.c_int(offsets[0]), C.c_int(offsets[1]), C.c_int(offsets[2]), C.c_uint32(direction), _p(fi), _p(u), _p(flags), C.c_uint64(t), C.c_uint8(flag), _p(p0), _p(p1), _p(p2), _p(bbu))

    def extract_fi(self, axis, t, bp, bm, fi):
        self.lib.orc_transfer_extract_fi(C.byref(self.g), C.c_uint32(axis), C.c_uint64(t), _p(bp), _p(bm), _p(fi))

    def insert_fi(self, axis, t, bp, bm, fi):
        self.lib.orc_transfer_insert_fi(C.byref(self.g), C.c_uint32(axis), C.c_uint64(t), _p(bp), _p(bm), _p(fi))

    def extract_ruf(self, axis, t, bp, bm, rho, u, flags):
        self.lib.orc_transfer_extract_rho_u_flags(C.byref(self.g), C.c_uint32(axis), C.c_uint64(t), _p(bp), _p(bm), _p(rho), _p(u), _p(flags))

    def insert_ruf(self, axis, t, bp, bm, rho, u, flags):
        self.lib.orc_transfer_insert_rho_u_flags(C.byref(self.g), C.c_uint32(axis), C.c_uint64(t), _p(bp), _p(bm), _p(rho), _p(u), _p(flags))


def ref_variant_name(Q, collision, storage, features):
    return f"q{Q}_{COLL_NAMES[collision]}_{STORAGE_NAMES[storage]}_f{features}"


def ref_available(Q, collision, storage, features):
    return os.path.exists(os.path.join(REF_DIR, f"libref_{ref_variant_name(Q, collision, storage, features)}.so"))


class RefBackend:
    """oracle/_ref: the reference's own kernel source compiled natively (built by oracle/ref/build_ref.py)"""
    name = "reference"

    def __init__(self, Q=19, collision=SRT, storage=FP32, features=0, threads=0):
        path = os.path.join(REF_DIR, f"libref_{ref_variant_name(Q, collision, storage, features)}.so")
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_float_to_half_custom.restype = C.c_uint16; L.ref_float_to_half_custom.argtypes = [C.c_float]
        L.ref_half_to_float_custom.restype = C.c_float; L.ref_half_to_float_custom.argtypes = [C.c_uint16]
        L.ref_set_threads(threads)
        assert L.ref_velocity_set() == Q
        self.Q, self.collision, self.storage, self.features = Q, collision, storage, features

    def set_grid(self, Nx, Ny, Nz, Dx, Dy, Dz, w):
        self.lib.ref_set_grid(Nx, Ny, Nz, Dx, Dy, Dz)
        self.lib.ref_set_w(C.c_float(w))

    def initialize(self, fi, rho, u, flags):
        self.lib.ref_initialize(_p(fi), _p(rho), _p(u), _p(flags))

    def stream_collide(self, fi, rho, u, flags, t, fx=0.0, fy=0.0, fz=0.0, F=None):
        if self.features & FORCE_FIELD: self.lib.ref_stream_collide_F(_p(fi), _p(rho), _p(u), _p(flags), C.c_uint64(t), C.c_float(fx), C.c_float(fy), C.c_float(fz), _p(F))
        else: self.lib.ref_stream_collide(_p(fi), _p(rho), _p(u), _p(flags), C.c_uint64(t), C.c_float(fx), C.c_float(fy), C.c_float(fz))

    def update_fields(self, fi, rho, u, flags, t, fx=0.0, fy=0.0, fz=0.0, F=None):
        if self.features & FORCE_FIELD: self.lib.ref_update_fields_F(_p(fi), _p(rho), _p(u), _p(flags), C.c_uint64(t), C.c_float(fx), C.c_float(fy), C.c_float(fz), _p(F))
        else: self.lib.ref_update_fields(_p(fi), _p(rho), _p(u), _p(flags), C.c_uint64(t), C.c_float(fx), C.c_float(fy), C.c_float(fz))

    def update_force_field(self, fi, flags, t, F):
        self.lib.ref_update_force_field(_p(fi), _p(flags), C.c_uint64(t), _p(F))

    def reset_force_field(self, F):
        self.lib.ref_reset_force_field(_p(F))

    def extract_F(self, axis, t, bp, bm, F):
        self.lib.ref_transfer_extract_F(C.c_uint32(axis), C.c_uint64(t), _p(bp), _p(bm), _p(F))

    def insert_F(self, axis, t, bp, bm, F):
        self.lib.ref_transfer_insert_F(C.c_uint32(axis), C.c_uint64(t), _p(bp), _p(bm), _p(F))

    def update_moving_boundaries(self, u, flags):
        self.lib.ref_update_moving_boundaries(_p(u), _p(flags))

    def voxelize_mesh(self, offsets, direction, fi, u, flags, t, flag, p0, p1, p2, bbu):
        self.lib.ref_set_offsets(C.c_int(offsets[0]), C.c_int(offsets[1]), C.c_int(offsets[2]))
        self.lib.ref_voxelize_mesh(C.c_uint32(direction), _p(fi), _p(u), _p(flags), C.c_uint64(t), C.c_uint8(flag), _p(p0), _p(p1), _p(p2), _p(bbu))

    def extract_fi(self, axis, t, bp, bm, fi):
        self.lib.ref_transfer_extract_fi(C.c_uint32(axis), C.c_uint64(t), _p(bp), _p(bm), _p(fi))

    def insert_fi(self, axis, t, bp, bm, fi):
        self.lib.ref_transfer_insert_fi(C.c_uint32(axis), C.c_uint64(t), _p(bp), _p(bm), _p(fi))

    def extract_ruf(self, axis, t, bp, bm, rho, u, flags):
        self.lib.ref_transfer_extract_rho_u_flags(C.c_uint32(axis), C.c_uint64(t), _p(bp), _p(bm), _p(rho), _p(u), _p(flags))

    def insert_ruf(self, axis, t, bp, bm, rho, u, flags):
        self.lib.ref_transfer_insert_rho_u_flags(C.c_uint32(axis), C.c_uint64(t), _p(bp), _p(bm), _p(rho), _p(u), _p(flags))


class Domain:
    def __init__(self, N, Q, storage, force_field=False):
        self.F = np.zeros(3 * N, dtype=np.float32) if force_field else None  # src/lbm.cpp:132
        self.fi = np.zeros(Q * N, dtype=ddf_dtype(storage))
        self.rho = np.ones(N, dtype=np.float32)          # src/lbm.cpp:124
        self.u = np.zeros(3 * N, dtype=np.float32)
        self.flags = np.zeros(N, dtype=np.uint8)
        self.buf_p = None
        self.buf_m = None


class HostSim:
    """Host sequencing of the reference (class LBM, src/lbm.cpp) over a kernel backend, D domains in one address space."""

    def __init__(self, backend, Nx, Ny, Nz, Dx=1, Dy=1, Dz=1, nu=1.0 / 6.0, w=None, fx=0.0, fy=0.0, fz=0.0):
        self.b = backend
        self.Nx, self.Ny, self.Nz = (Nx // Dx) * Dx, (Ny // Dy) * Dy, (Nz // Dz) * Dz   # src/lbm.cpp:722-724
        self.Dx, self.Dy, self.Dz = Dx, Dy, Dz
        self.D = Dx * Dy * Dz
        self.Hx, self.Hy, self.Hz = int(Dx > 1), int(Dy > 1), int(Dz > 1)
        self.lNx, self.lNy, self.lNz = self.Nx // Dx + 2 * self.Hx, self.Ny // Dy + 2 * self.Hy, self.Nz // Dz + 2 * self.Hz
        self.lN = self.lNx * self.lNy * self.lNz
        self.w = w if w is not None else OracleBackend().w_from_nu(nu)
        self.f = (fx, fy, fz)
        self.t = 0
        self.initialized = False
        self.ff = bool(backend.features & FORCE_FIELD)
        self.t_last_force_field = None
        self.dom = [Domain(self.lN, backend.Q, backend.storage, self.ff) for _ in range(self.D)]
        if self.D > 1:
            A = max([self.lNy * self.lNz] * self.Hx + [self.lNz * self.lNx] * self.Hy + [self.lNx * self.lNy] * self.Hz)
            T = 5 if backend.Q == 19 else 9
            per = max(T * (4 if backend.storage == FP32 else 2), 17)   # src/lbm.cpp:1314-1315 (17 >= the 12 bytes of F)
            for d in self.dom:
                d.buf_p = np.zeros(A * per, dtype=np.uint8)
                d.buf_m = np.zeros(A * per, dtype=np.uint8)
        self.b.set_grid(self.lNx, self.lNy, self.lNz, Dx, Dy, Dz, self.w)

    # ---- global <-> domain mapping, src/lbm.hpp:265-288 ----
    def _views(self, name, dim=None):
        """list of (domain interior view, global slices)"""
        out = []
        for d in range(self.D):
            x, y, z = (d % (self.Dx * self.Dy)) % self.Dx, (d % (self.Dx * self.Dy)) // self.Dx, d // (self.Dx * self.Dy)
            arr = getattr(self.dom[d], name)
            if dim is not None:
                arr = arr[dim * self.lN:(dim + 1) * self.lN]
            a3 = arr.reshape(self.lNz, self.lNy, self.lNx)
            inner = a3[self.Hz:self.lNz - self.Hz, self.Hy:self.lNy - self.Hy, self.Hx:self.lNx - self.Hx]
            nx, ny, nz = self.Nx // self.Dx, self.Ny // self.Dy, self.Nz // self.Dz
            out.append((inner, (slice(z * nz, (z + 1) * nz), slice(y * ny, (y + 1) * ny), slice(x * nx, (x + 1) * nx))))
        return out

    def set_global(self, name, values, dim=None):
        """values: array shaped (Nz,Ny,Nx) on the halo-free global grid"""
        for inner, sl in self._views(name, dim):
            inner[...] = values[sl]

    def get_global(self, name, dim=None):
        proto = getattr(self.dom[0], name)
        out = np.zeros((self.Nz, self.Ny, self.Nx), dtype=proto.dtype)
        for inner, sl in self._views(name, dim):
            out[sl] = inner
        return out

    # ---- src/lbm.cpp:1343-1390 ----
    def _communicate(self, field):
        for axis, Dn in enumerate((self.Dx, self.Dy, self.Dz)):
            if Dn <= 1:
                continue
            for d in self.dom:
                if field == "fi":
                    self.b.extract_fi(axis, self.t, d.buf_p, d.buf_m, d.fi)
                elif field == "F":
                    self.b.extract_F(axis, self.t, d.buf_p, d.buf_m, d.F)
                else:
                    self.b.extract_ruf(axis, self.t, d.buf_p, d.buf_m, d.rho, d.u, d.flags)
            for di in range(self.D):
                x, y, z = (di % (self.Dx * self.Dy)) % self.Dx, (di % (self.Dx * self.Dy)) // self.Dx, di // (self.Dx * self.Dy)
                c = [x, y, z]
                c[axis] = (c[axis] + 1) % Dn
                dn = c[0] + (c[1] + c[2] * self.Dy) * self.Dx
                self.dom[di].buf_p, self.dom[dn].buf_m = self.dom[dn].buf_m, self.dom[di].buf_p
            for d in self.dom:
                if field == "fi":
                    self.b.insert_fi(axis, self.t, d.buf_p, d.buf_m, d.fi)
                elif field == "F":
                    self.b.insert_F(axis, self.t, d.buf_p, d.buf_m, d.F)
                else:
                    self.b.insert_ruf(axis, self.t, d.buf_p, d.buf_m, d.rho, d.u, d.flags)

    def initialize(self):  # src/lbm.cpp:881-922
        if self.ff: self._communicate("F")  # :889-892
        self.t = 1
        self._communicate("ruf")
        for d in self.dom:
            self.b.initialize(d.fi, d.rho, d.u, d.flags)
        self._communicate("ruf")
        self._communicate("fi")
        self.t = 0
        self.initialized = True

    def run(self, steps):  # src/lbm.cpp:924-975
        if not self.initialized:
            self.initialize()
        for _ in range(steps):
            for d in self.dom:
                self.b.stream_collide(d.fi, d.rho, d.u, d.flags, self.t, *self.f, **({"F": d.F} if self.ff else {}))
            self._communicate("fi")
            self.t += 1

    def update_moving_boundaries(self):  # src/lbm.cpp:1018-1027 (the flags halo travels with rho and u here; both are unchanged copies)
        for d in self.dom:
            self.b.update_moving_boundaries(d.u, d.flags)
        self._communicate("ruf")

    def voxelize_mesh(self, mesh, flag=TYPE_S, rotation_center=None, linear_velocity=(0.0, 0.0, 0.0), rotational_velocity=(0.0, 0.0, 0.0)):
        """LBM::voxelize_mesh_on_device (src/lbm.cpp:1074-1089) over LBM_Domain::voxelize_mesh_on_device (:275-322) on every domain"""
        bbu, direction = voxelize_parameters(mesh, rotation_center if rotation_center is not None else mesh.center, linear_velocity, rotational_velocity)
        nx, ny, nz = self.Nx // self.Dx, self.Ny // self.Dy, self.Nz // self.Dz
        for d in range(self.D):
            x, y, z = (d % (self.Dx * self.Dy)) % self.Dx, (d % (self.Dx * self.Dy)) // self.Dx, d // (self.Dx * self.Dy)
            off = (x * nx - self.Hx, y * ny - self.Hy, z * nz - self.Hz)  # src/lbm.cpp:733
            dom = self.dom[d]
            self.b.voxelize_mesh(off, direction, dom.fi, dom.u, dom.flags, self.t + 1, flag, mesh.p0, mesh.p1, mesh.p2, bbu)
        if (self.b.features & MOVING_BOUNDARIES) and (flag & (TYPE_S | TYPE_E)) == TYPE_S and (any(v != 0.0 for v in linear_velocity) or any(v != 0.0 for v in rotational_velocity)):
            self.update_moving_boundaries()

    def update_fields(self):  # src/lbm.cpp:977-980
        for d in self.dom:
            self.b.update_fields(d.fi, d.rho, d.u, d.flags, self.t, *self.f, **({"F": d.F} if self.ff else {}))

    # ---- FORCE_FIELD, src/lbm.cpp:206-239,986-1016 ----
    def update_force_field(self):
        if self.t != self.t_last_force_field:  # :207-211
            for d in self.dom:
                self.b.update_force_field(d.fi, d.flags, self.t, d.F)
            self.t_last_force_field = self.t

    def object_sum(self, kind, marker, center=(0.0, 0.0, 0.0), group=64):
        """LBM::object_center_of_mass / object_force / object_torque: per-domain sums added in domain order (:992-1015); oracle backend only"""
        if kind != 0: self.update_force_field()
        parts = [self.b.object_sum(kind, d.F, d.flags, marker, center, group) for d in self.dom]
        tot = np.zeros(3, np.float32); cells = 0
        for p in parts:
            tot = (tot + p[:3]).astype(np.float32); cells += int(p[3:4].view(np.uint32)[0])
        return (tot / np.float32(cells)).astype(np.float32) if kind == 0 else tot

    def fields(self):
        """(rho, ux, uy, uz, flags) on the global grid after update_fields (what lbm.u.read_from_device() returns)"""
        if not (self.b.features & UPDATE_FIELDS):  # src/lbm.hpp:390-393: with UPDATE_FIELDS the fields are already current
            self.update_fields()
        return (self.get_global("rho"), self.get_global("u", 0), self.get_global("u", 1), self.get_global("u", 2), self.get_global("flags"))


def hash01(n, seed):
    """deterministic integer hash -> float32 in [0,1); identical in numpy, C and CUDA (SURVEY section 8d, C1 parity IC)"""
    x = (np.asarray(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(30); x = (x * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(27); x = (x * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(31)
    return ((x >> np.uint64(40)).astype(np.float32) / np.float32(16777216.0)).astype(np.float32)


def scenario(Nx, Ny, Nz, seed=1, solid_frac=0.08, eq_frac=0.0, amp_rho=0.01, amp_u=0.05):
    """perturbed initial condition + random obstacles on the global grid, shaped (Nz,Ny,Nx)"""
    n = np.arange(Nx * Ny * Nz, dtype=np.uint64).reshape(Nz, Ny, Nx)
    with np.errstate(over="ignore"):
        rho = (np.float32(1.0) + np.float32(amp_rho) * (hash01(n, seed) - np.float32(0.5))).astype(np.float32)
        u = [(np.float32(amp_u) * (hash01(n, seed + 1 + a) - np.float32(0.5))).astype(np.float32) for a in range(3)]
        r = hash01(n, seed + 7)
    flags = np.zeros((Nz, Ny, Nx), dtype=np.uint8)
    flags[r < solid_frac] = TYPE_S
    flags[(r >= solid_frac) & (r < solid_frac + eq_frac)] = TYPE_E
    return rho, u, flags


def load_scenario(sim, rho, u, flags):
    sim.set_global("rho", rho)
    for a in range(3):
        sim.set_global("u", u[a], a)
    sim.set_global("flags", flags)


# ---- triangle meshes: binary STL (src/utilities.hpp:4530-4581) and the parameter block of the voxeliser (src/lbm.cpp:275-322) ----
class Mesh:
    """p0, p1, p2: float32 arrays (triangles, 3), C-contiguous; center, pmin, pmax: float32[3] (src/utilities.hpp:4425-4528)"""

    def __init__(self, p0, p1, p2, center):
        self.p0, self.p1, self.p2 = (np.ascontiguousarray(a, dtype=np.float32) for a in (p0, p1, p2))
        self.center = np.asarray(center, dtype=np.float32)
        self.find_bounds()

    @property
    def triangle_number(self): return self.p0.shape[0]

    def find_bounds(self):
        allp = np.concatenate([self.p0, self.p1, self.p2])
        self.pmin, self.pmax = allp.min(axis=0), allp.max(axis=0)


def read_stl(path, box_size, center, size, rotation=None, reposition=True):
    """read_stl_raw, src/utilities.hpp:4530-4571: all arithmetic in binary32 in the reference's order center+scale*(offset+p)"""
    raw = open(path, "rb").read()
    n = int(np.frombuffer(raw, np.uint32, 1, 80)[0])
    assert n > 0 and len(raw) == 84 + 50 * n, "corrupt or non-binary STL"
    rec = np.frombuffer(raw, np.uint8, 50 * n, 84).reshape(n, 50)[:, :48].copy().view(np.float32).reshape(n, 12)
    p = [rec[:, 3:6].copy(), rec[:, 6:9].copy(), rec[:, 9:12].copy()]
    if rotation is not None:
        R = np.asarray(rotation, dtype=np.float32)
        p = [np.stack([(R[r, 0] * a[:, 0] + R[r, 1] * a[:, 1] + R[r, 2] * a[:, 2]).astype(np.float32) for r in range(3)], axis=1) for a in p]
    m = Mesh(*p, center)
    f32 = np.float32
    ext = (m.pmax - m.pmin).astype(f32)
    box = np.asarray(box_size, dtype=f32)
    if size == 0.0: scale = f32(min(box[0] / ext[0], box[1] / ext[1], box[2] / ext[2]))
    elif size > 0.0: scale = f32(f32(size) / max(ext))
    else: scale = f32(-size)
    offset = (f32(-0.5) * (m.pmin + m.pmax)).astype(f32) if reposition else np.zeros(3, f32)
    c = np.asarray(center, dtype=f32)
    out = [(c + scale * (offset + a)).astype(f32) for a in p]
    return Mesh(*out, c)


def write_stl(path, p0, p1, p2):
    n = p0.shape[0]
    rec = np.zeros((n, 50), np.uint8)
    body = np.zeros((n, 12), np.float32)
    body[:, 3:6], body[:, 6:9], body[:, 9:12] = p0, p1, p2
    rec[:, :48] = body.view(np.uint8).reshape(n, 48)
    with open(path, "wb") as f:
        f.write(b"fx3d test mesh".ljust(80, b" ")); f.write(np.uint32(n).tobytes()); f.write(rec.tobytes())


def voxelize_parameters(mesh, rotation_center, linear_velocity, rotational_velocity):
    """the 16-float block and the ray direction LBM_Domain::voxelize_mesh_on_device chooses (src/lbm.cpp:279-321)"""
    f32 = np.float32
    bbu = np.zeros(16, f32)
    bbu[0] = np.array([mesh.triangle_number], np.uint32).view(f32)[0]
    bbu[1:4] = mesh.pmin - f32(2.0); bbu[4:7] = mesh.pmax + f32(2.0)
    bbu[7:10] = np.asarray(rotation_center, f32); bbu[10:13] = np.asarray(linear_velocity, f32); bbu[13:16] = np.asarray(rotational_velocity, f32)
    x0, y0, z0, x1, y1, z1 = bbu[1:7]
    rot = bbu[13:16]
    if float(np.sqrt(f32(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2]))) == 0.0:  # direction of the smallest bounding-box cross-section
        v = [f32((y1 - y0) * (z1 - z0)), f32((z1 - z0) * (x1 - x0)), f32((x1 - x0) * (y1 - y0))]
        direction = 0
        for i in (1, 2):
            if v[i] < v[direction]: direction = i
    else:  # direction closest to the rotation axis
        v = [abs(float(r)) for r in rot]
        direction = 0
        for i in (1, 2):
            if v[i] > v[direction]: direction = i
    return bbu, direction


def torus_mesh(R=1.0, r=0.4, nu=24, nv=12):
    """a closed, non-convex test surface: torus around the z axis, nu x nv quads split into triangles"""
    a = np.linspace(0, 2 * np.pi, nu, endpoint=False); b = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    A, B = np.meshgrid(a, b, indexing="ij")
    P = np.stack([(R + r * np.cos(B)) * np.cos(A), (R + r * np.cos(B)) * np.sin(A), r * np.sin(B)], axis=-1).astype(np.float32)
    p0, p1, p2 = [], [], []
    for i in range(nu):
        for j in range(nv):
            q00, q10, q01, q11 = P[i, j], P[(i + 1) % nu, j], P[i, (j + 1) % nv], P[(i + 1) % nu, (j + 1) % nv]
            p0 += [q00, q10]; p1 += [q10, q11]; p2 += [q01, q01]
    return np.array(p0, np.float32), np.array(p1, np.float32), np.array(p2, np.float32)
