"""Python mirror of the reference's host surface for the LBM hot path: class LBM / LBM_Domain / Memory_Container
(FluidX3D v3.7 src/lbm.hpp:20-204,208-611; src/lbm.cpp:96-191,700-980,1308-1390), driving libfx3d_cuda.so through
its C ABI (include/fx3d.h). Same names, argument meaning and sequencing as the reference; what the reference fixes at
compile time in src/defines.hpp (D3Q19/D3Q27, SRT/TRT, FP16S/FP16C, VOLUME_FORCE, EQUILIBRIUM_BOUNDARIES, UPDATE_FIELDS)
is a constructor keyword here.

Two multi-GPU modes share one code path:
  * single process, D domains on one or several GPUs (the reference's model, src/lbm.cpp:721-734); all domains may share
    one GPU, which the reference also allows ("Using single fastest device for all domains", src/lbm.cpp:663-664)
  * one process per GPU (torchrun): pass `comm`; this process then owns domain `comm.rank` and maps its neighbours'
    buffers through CUDA IPC.
In both, LBM::communicate_field's device->host->device round trip (src/lbm.cpp:1355-1383) is replaced by direct peer
pulls (fx3d_exchange_*) ordered by device-side rendezvous counters; no step of run() synchronises with the host.

The C++ twin of this file (what setup.cpp scenes compile against) is fluidx3d_b200/host/lbm.hpp.
"""
import ctypes as C
import numpy as np
from . import capi
from .capi import (FP32, FP16S, FP16C, SRT, TRT, VOLUME_FORCE, EQUILIBRIUM_BOUNDARIES, UPDATE_FIELDS, SUBGRID, MOVING_BOUNDARIES, FORCE_FIELD, TYPE_S, TYPE_E,
                   REGION_ALL, REGION_SHELL, REGION_INTERIOR, Fx3dError)

max_ulong = 2 ** 64 - 1


VERBOSE = True  # module switch for the console messages below (tests turn it off)


def print_warning(msg):  # src/utilities.hpp:4066-4073
    if VERBOSE:
        print(f"| Warning: {msg}")


def _domain_xyz(d, Dx, Dy):  # d = x+(y+z*Dy)*Dx, src/lbm.cpp:732
    return (d % (Dx * Dy)) % Dx, (d % (Dx * Dy)) // Dx, d // (Dx * Dy)


class Memory:
    """Memory<T> (src/opencl.hpp:342-614) for one domain: page-locked host array + device buffer of N*dims elements."""

    def __init__(self, lib, device, N, dims, dtype, alloc_host=True, value=0):
        self.lib, self.device, self.N, self.d = lib, device, int(N), int(dims)
        self.dtype = np.dtype(dtype)
        self.nbytes = self.N * self.d * self.dtype.itemsize
        p = C.c_void_p()
        lib.malloc(device, self.nbytes, C.byref(p))
        self.device_ptr = p.value
        self.host_ptr, self.host = None, None
        if alloc_host:
            h = C.c_void_p()
            lib.host_alloc(self.nbytes, C.byref(h))
            self.host_ptr = h.value
            self.host = np.frombuffer((C.c_char * self.nbytes).from_address(h.value), dtype=self.dtype)
            self.host[...] = value
        self.initial = value

    def length(self): return self.N
    def dimensions(self): return self.d
    def range(self): return self.N * self.d
    def capacity(self): return self.nbytes

    def enqueue_write_to_device(self, stream):
        self.lib.memcpy_h2d(self.device, self.device_ptr, self.host_ptr, self.nbytes, stream, 0)

    def enqueue_read_from_device(self, stream):
        self.lib.memcpy_d2h(self.device, self.host_ptr, self.device_ptr, self.nbytes, stream, 0)

    def free(self):
        if self.device_ptr:
            self.lib.free(self.device, self.device_ptr); self.device_ptr = None
        if self.host_ptr:
            self.host = None
            self.lib.host_free(self.host_ptr); self.host_ptr = None


class LBM_Domain:
    """One device's share of the lattice (src/lbm.hpp:20-204): local size includes the halo layers."""

    def __init__(self, lib, device, stream, Nx, Ny, Nz, Dx, Dy, Dz, Ox, Oy, Oz, nu, fx, fy, fz, velocity_set, collision, storage, features, host_fields=True):
        self.lib, self.device, self.stream = lib, device, stream
        self.Nx, self.Ny, self.Nz, self.Dx, self.Dy, self.Dz, self.Ox, self.Oy, self.Oz = Nx, Ny, Nz, Dx, Dy, Dz, Ox, Oy, Oz
        self.nu, self.fx, self.fy, self.fz = nu, fx, fy, fz
        self.t = 0
        self.t_last_update_fields = max_ulong
        self.t_last_force_field = max_ulong
        self.velocity_set, self.features = velocity_set, features
        N = self.get_N()
        self.lat = capi.Lattice(device, Nx, Ny, Nz, Dx, Dy, Dz, velocity_set, collision, storage, features, lib.relaxation_rate(C.c_float(nu)), None, None, None, None, None)
        # allocate(), src/lbm.cpp:121-129: fi device only; rho (=1), u, flags host+device
        fi_bytes = lib.fi_bytes(C.byref(self.lat))
        self.fi = Memory(lib, device, fi_bytes, 1, np.uint8, alloc_host=False)
        self.rho = Memory(lib, device, N, 1, np.float32, alloc_host=host_fields, value=1.0)
        self.u = Memory(lib, device, N, 3, np.float32, alloc_host=host_fields)
        self.flags = Memory(lib, device, N, 1, np.uint8, alloc_host=host_fields)
        if not host_fields:  # benchmark path: default fields (rho=1, u=0, flags=0) are produced on the device
            lib.fill_f32(device, self.rho.device_ptr, C.c_float(1.0), N, stream)
        self.lat.fi, self.lat.rho, self.lat.u, self.lat.flags = self.fi.device_ptr, self.rho.device_ptr, self.u.device_ptr, self.flags.device_ptr
        self.F, self.object_sum, self.object_scratch = None, None, None
        if features & FORCE_FIELD:  # src/lbm.cpp:131-141: F (host+device) and object_sum (x, y, z, cell count)
            self.F = Memory(lib, device, N, 3, np.float32, alloc_host=True)
            self.object_sum = Memory(lib, device, 1, 4, np.float32, alloc_host=True)
            self.object_scratch = Memory(lib, device, lib.object_scratch_bytes(C.byref(self.lat)), 1, np.uint8, alloc_host=False)
            self.lat.F = self.F.device_ptr
        self.sync_array, self.xfer, self.xfer_bytes = None, None, 0
        if Dx * Dy * Dz > 1:  # allocate_transfer(), src/lbm.cpp:1308-1337: the rendezvous counters ...
            p = C.c_void_p()
            lib.malloc(device, 64 * 8, C.byref(p))
            self.sync_array = p.value
        if Dx > 1:  # ... and, for x faces only (one element per row: strided), a pair of linear staging buffers [+x | -x]
            self.xfer_bytes = (lib.transfer_bytes(C.byref(self.lat)) + 255) // 256 * 256
            p = C.c_void_p()
            lib.malloc(device, 2 * self.xfer_bytes, C.byref(p))
            self.xfer = p.value

    def get_N(self): return self.Nx * self.Ny * self.Nz
    def get_D(self): return self.Dx * self.Dy * self.Dz
    def get_t(self): return self.t
    def get_tau(self): return 3.0 * self.nu + 0.5
    def set_f(self, fx, fy, fz): self.fx, self.fy, self.fz = fx, fy, fz

    def enqueue_initialize(self):  # src/lbm.cpp:178-180
        self.lib.initialize(C.byref(self.lat), self.stream)

    def enqueue_stream_collide(self, region=REGION_ALL, stream=None):  # src/lbm.cpp:181-183: t, fx, fy, fz are per-launch arguments
        self.lib.stream_collide(C.byref(self.lat), self.t, self.fx, self.fy, self.fz, region, stream if stream is not None else self.stream)

    def enqueue_update_moving_boundaries(self):  # src/lbm.cpp:241-243
        self.lib.update_moving_boundaries(C.byref(self.lat), self.stream)

    # ---- FORCE_FIELD, src/lbm.cpp:206-239 ----
    def enqueue_update_force_field(self):
        if self.t != self.t_last_force_field:  # only if the time step has changed since the last update
            self.lib.update_force_field(C.byref(self.lat), self.t, self.stream)
            self.t_last_force_field = self.t

    def enqueue_object_center_of_mass(self, flag_marker):
        self.lib.object_center_of_mass(C.byref(self.lat), flag_marker, self.object_sum.device_ptr, self.object_scratch.device_ptr, self.stream)
        self.object_sum.enqueue_read_from_device(self.stream)

    def enqueue_object_force(self, flag_marker):
        self.enqueue_update_force_field()
        self.lib.object_force(C.byref(self.lat), flag_marker, self.object_sum.device_ptr, self.object_scratch.device_ptr, self.stream)
        self.object_sum.enqueue_read_from_device(self.stream)

    def enqueue_object_torque(self, rotation_center, flag_marker):
        self.enqueue_update_force_field()
        self.lib.object_torque(C.byref(self.lat), flag_marker, C.c_float(rotation_center[0]), C.c_float(rotation_center[1]), C.c_float(rotation_center[2]),
                               self.object_sum.device_ptr, self.object_scratch.device_ptr, self.stream)
        self.object_sum.enqueue_read_from_device(self.stream)

    def enqueue_update_fields(self):  # src/lbm.cpp:184-191
        if not (self.features & UPDATE_FIELDS) and self.t != self.t_last_update_fields:
            self.lib.update_fields(C.byref(self.lat), self.t, self.fx, self.fy, self.fz, self.stream)
            self.t_last_update_fields = self.t

    def increment_time_step(self, steps=1):  # src/lbm.cpp:255-260
        self.t += steps
        if self.features & UPDATE_FIELDS:
            self.t_last_update_fields = self.t

    def reset_time_step(self):
        self.t = 0
        if self.features & UPDATE_FIELDS:
            self.t_last_update_fields = self.t

    def finish_queue(self):
        self.lib.stream_sync(self.device, self.stream)

    def free(self):
        for m in (self.fi, self.rho, self.u, self.flags, self.F, self.object_sum, self.object_scratch):
            if m is not None: m.free()
        if self.sync_array:
            self.lib.free(self.device, self.sync_array); self.sync_array = None
        if self.xfer:
            self.lib.free(self.device, self.xfer); self.xfer = None


class Memory_Container:
    """Stitches the per-domain host buffers into one global index space (src/lbm.hpp:239-408). `lbm.rho[n]`,
    `lbm.u.x[n]` take a global linear index n = x+(y+z*Ny)*Nx on the halo-free grid; whole-array access through
    get_global()/set_global() (shape (Nz,Ny,Nx)) is the vectorised equivalent of a parallel_for over n."""

    class Pointer:
        def __init__(self, mc, dim): self.mc, self.dim = mc, dim
        def __getitem__(self, i): return self.mc._ref(i, self.dim)[0][self.mc._ref(i, self.dim)[1]]
        def __setitem__(self, i, v):
            arr, k = self.mc._ref(i, self.dim); arr[k] = v

    def __init__(self, lbm, name, dims):
        self.lbm, self.name, self.d = lbm, name, dims
        self.N = lbm.get_N()
        self.x = Memory_Container.Pointer(self, 0)
        if dims > 1: self.y = Memory_Container.Pointer(self, 1)
        if dims > 2: self.z = Memory_Container.Pointer(self, 2)

    def _mem(self, dom): return getattr(dom, self.name)
    def length(self): return self.N
    def dimensions(self): return self.d
    def range(self): return self.N * self.d

    def _ref(self, i, dim=0):  # reference(i, dimension), src/lbm.hpp:265-288
        L = self.lbm
        gi = i % self.N
        dim = max(i // self.N, dim)
        t = gi % (L.Nx * L.Ny)
        x, y, z = t % L.Nx, t // L.Nx, gi // (L.Nx * L.Ny)
        nx, ny, nz = L.Nx // L.Dx, L.Ny // L.Dy, L.Nz // L.Dz
        d = x // nx + (y // ny + (z // nz) * L.Dy) * L.Dx
        dom = L.domain_or_none(d)
        if dom is None:
            raise IndexError(f"cell {i} belongs to domain {d}, which another process owns")
        li = (x % nx + L.Hx) + ((y % ny + L.Hy) + (z % nz + L.Hz) * dom.Ny) * dom.Nx
        return self._mem(dom).host, li + dim * dom.get_N()

    def __getitem__(self, i):
        arr, k = self._ref(i); return arr[k]

    def __setitem__(self, i, v):
        arr, k = self._ref(i); arr[k] = v

    def _views(self, dim):
        L = self.lbm
        nx, ny, nz = L.Nx // L.Dx, L.Ny // L.Dy, L.Nz // L.Dz
        for d, dom in L.local_domains():
            x, y, z = _domain_xyz(d, L.Dx, L.Dy)
            N = dom.get_N()
            a3 = self._mem(dom).host[dim * N:(dim + 1) * N].reshape(dom.Nz, dom.Ny, dom.Nx)
            yield a3[L.Hz:dom.Nz - L.Hz, L.Hy:dom.Ny - L.Hy, L.Hx:dom.Nx - L.Hx], (slice(z * nz, (z + 1) * nz), slice(y * ny, (y + 1) * ny), slice(x * nx, (x + 1) * nx))

    def set_global(self, values, dim=0):
        """values: array of shape (Nz,Ny,Nx) on the global grid, or a scalar"""
        for inner, sl in self._views(dim):
            inner[...] = values[sl] if isinstance(values, np.ndarray) else values

    def get_global(self, dim=0):
        """global (Nz,Ny,Nx) array assembled from the host buffers of the domains this process owns (others stay 0)"""
        out = np.zeros((self.lbm.Nz, self.lbm.Ny, self.lbm.Nx), dtype=self._mem(self.lbm.local_domains()[0][1]).dtype)
        for inner, sl in self._views(dim):
            out[sl] = inner
        return out

    def reset(self, value=0):
        for _, dom in self.lbm.local_domains():
            self._mem(dom).host[...] = value

    def read_from_device(self):  # src/lbm.hpp:390-396
        L = self.lbm
        if not (L.features & UPDATE_FIELDS) and L.initialized:
            for _, dom in L.local_domains(): dom.enqueue_update_fields()
        for _, dom in L.local_domains(): self._mem(dom).enqueue_read_from_device(dom.stream)
        for _, dom in L.local_domains(): dom.finish_queue()

    def write_to_device(self):  # src/lbm.hpp:397-400
        for _, dom in self.lbm.local_domains(): self._mem(dom).enqueue_write_to_device(dom.stream)
        for _, dom in self.lbm.local_domains(): dom.finish_queue()


class LBM:
    """class LBM (src/lbm.hpp:208-611). LBM(Nx,Ny,Nz,nu,...) or LBM(Nx,Ny,Nz,Dx,Dy,Dz,nu,...) as in the reference
    (src/lbm.hpp:428-435): pass Dx,Dy,Dz by keyword."""

    def __init__(self, Nx, Ny, Nz, nu, fx=0.0, fy=0.0, fz=0.0, *, Dx=1, Dy=1, Dz=1, velocity_set=19, collision=SRT, storage=FP32,
                 features=0, devices=None, comm=None, lib=None, host_fields=True, overlap=False, benchmark=False, fuse_halo=True):
        self.lib = lib or capi.lib()
        NDx, NDy, NDz = (Nx // Dx) * Dx if Dx else 0, (Ny // Dy) * Dy if Dy else 0, (Nz // Dz) * Dz if Dz else 0
        self._sanity_checks_constructor(Nx, Ny, Nz, Dx, Dy, Dz, nu, fx, fy, fz, velocity_set, collision, storage, features)
        if (NDx, NDy, NDz) != (Nx, Ny, Nz):  # src/lbm.cpp:722-723
            print_warning(f"LBM grid ({Nx}x{Ny}x{Nz}) is not equally divisible in domains ({Dx}x{Dy}x{Dz}). Changing resolution to ({NDx}x{NDy}x{NDz}).")
        self.Nx, self.Ny, self.Nz, self.Dx, self.Dy, self.Dz = NDx, NDy, NDz, Dx, Dy, Dz
        self.Hx, self.Hy, self.Hz = int(Dx > 1), int(Dy > 1), int(Dz > 1)
        self.velocity_set, self.collision, self.storage, self.features = velocity_set, collision, storage, features
        self.initialized = False
        self.comm, self.overlap, self.benchmark, self.fuse_halo = comm, overlap, benchmark, fuse_halo
        self._fused = {}
        D = self.get_D()
        if comm is not None and comm.world_size != D:
            raise ValueError(f"{D} domains need {D} processes, got {comm.world_size}")
        owned = [comm.rank] if comm is not None else list(range(D))
        if devices is None:  # smart_device_selection, src/lbm.cpp:655-698: D identical devices, else one device for all domains
            ndev = self.lib.num_devices()
            if comm is not None:
                devices = {comm.rank: comm.local_rank % ndev}
            elif ndev >= D:
                devices = {d: d for d in range(D)}
            else:
                if D > 1: print_warning("Not enough devices of the same type available. Using single fastest device for all domains.")
                devices = {d: 0 for d in range(D)}
        elif not isinstance(devices, dict):
            devices = {d: devices[i] for i, d in enumerate(owned)}
        self._streams, self._streams2, self._events = {}, {}, {}
        self.lbm_domain = {}
        lx, ly, lz = NDx // Dx + 2 * self.Hx, NDy // Dy + 2 * self.Hy, NDz // Dz + 2 * self.Hz
        for d in owned:
            dev = devices[d]
            if dev not in self._streams:  # one in-order stream per physical device, shared by the domains that live on it
                s = C.c_void_p(); self.lib.stream_create(dev, C.byref(s)); self._streams[dev] = s.value
            x, y, z = _domain_xyz(d, Dx, Dy)
            self.lbm_domain[d] = LBM_Domain(self.lib, dev, self._streams[dev], lx, ly, lz, Dx, Dy, Dz,
                                            x * NDx // Dx - self.Hx, y * NDy // Dy - self.Hy, z * NDz // Dz - self.Hz,
                                            nu, fx, fy, fz, velocity_set, collision, storage, features, host_fields)
        if overlap and D > 1:  # second in-order stream per device for the interior cells, and the two events that order it with the first
            for dev in self._streams:
                s2 = C.c_void_p(); self.lib.stream_create(dev, C.byref(s2)); self._streams2[dev] = s2.value
                e1, e2 = C.c_void_p(), C.c_void_p()
                self.lib.event_create(dev, C.byref(e1)); self.lib.event_create(dev, C.byref(e2))
                self._events[dev] = (e1.value, e2.value)
        self.host_fields = host_fields
        self.rho = Memory_Container(self, "rho", 1)
        self.u = Memory_Container(self, "u", 3)
        self.flags = Memory_Container(self, "flags", 1)
        if features & FORCE_FIELD: self.F = Memory_Container(self, "F", 3)  # src/lbm.hpp:412-414
        self._seq = 0
        self._peers = None
        if D > 1:
            self._connect_peers()

    # ---- checks, src/lbm.cpp:783-838 ----
    def _sanity_checks_constructor(self, Nx, Ny, Nz, Dx, Dy, Dz, nu, fx, fy, fz, velocity_set, collision, storage, features):
        if Nx * Ny * Nz == 0: raise ValueError(f"Grid point number is 0: {Nx}x{Ny}x{Nz} = 0.")
        if Dx * Dy * Dz == 0: raise ValueError(f"You specified 0 LBM grid domains ({Dx}x{Dy}x{Dz}). There has to be at least 1 domain in every direction.")
        if Dx * Dy * Dz > 63: raise ValueError("at most 63 domains")
        if nu == 0.0: raise ValueError("Viscosity cannot be 0.")
        if nu < 0.0: raise ValueError("Viscosity cannot be negative.")
        if velocity_set not in (19, 27): raise ValueError("velocity_set must be 19 (D3Q19) or 27 (D3Q27)")
        if collision not in (SRT, TRT): raise ValueError("No LBM collision operator selected: collision must be SRT or TRT")
        if storage not in (FP32, FP16S, FP16C): raise ValueError("storage must be FP32, FP16S or FP16C")
        if not (features & VOLUME_FORCE) and (fx != 0.0 or fy != 0.0 or fz != 0.0):
            raise ValueError("Volume force is set in LBM constructor, but VOLUME_FORCE is not enabled.")
        if (features & VOLUME_FORCE) and fx == 0.0 and fy == 0.0 and fz == 0.0:
            print_warning("The VOLUME_FORCE extension is enabled but the volume force in LBM constructor is set to zero.")

    def _sanity_checks_initialization(self):  # src/lbm.cpp:840-879 (flags scan; skipped under BENCHMARK, :882-884)
        used_e, moving = False, False
        for _, dom in self.local_domains():
            bo = dom.flags.host & (TYPE_S | TYPE_E)
            used_e |= bool(np.any(bo == TYPE_E))
            N = dom.get_N()
            solid = bo == TYPE_S
            if np.any(solid):
                moving |= bool(np.any((dom.u.host[:N][solid] != 0) | (dom.u.host[N:2 * N][solid] != 0) | (dom.u.host[2 * N:][solid] != 0)))
        if moving and not (self.features & MOVING_BOUNDARIES): print_warning("Some boundary cells have non-zero velocity, but MOVING_BOUNDARIES is not enabled.")
        if used_e and not (self.features & EQUILIBRIUM_BOUNDARIES):
            raise ValueError("Some cells are set as equilibrium boundaries with the TYPE_E flag, but EQUILIBRIUM_BOUNDARIES is not enabled.")
        if not used_e and (self.features & EQUILIBRIUM_BOUNDARIES):
            print_warning("The EQUILIBRIUM_BOUNDARIES extension is enabled but no equilibrium boundary cells (TYPE_E flag) are placed in the simulation box.")

    # ---- domains ----
    def local_domains(self): return sorted(self.lbm_domain.items())
    def domain_or_none(self, d): return self.lbm_domain.get(d)

    def _neighbour(self, d, axis, sign):
        c = list(_domain_xyz(d, self.Dx, self.Dy)); Dn = (self.Dx, self.Dy, self.Dz)[axis]
        c[axis] = (c[axis] + sign) % Dn
        return c[0] + (c[1] + c[2] * self.Dy) * self.Dx

    def _neighbour_yz(self, d, dy, dz):
        x, y, z = _domain_xyz(d, self.Dx, self.Dy)
        return x + ((y + dy) % self.Dy + ((z + dz) % self.Dz) * self.Dy) * self.Dx

    def _sync_neighbours(self, d):
        """domains whose kernels read or write this domain's memory within a step: the face neighbours, plus -- with the y/z halo
        delivery fused into stream_collide -- the diagonal neighbours in the y-z plane"""
        nb = {self._neighbour(d, a, s) for a in range(3) for s in (1, -1) if (self.Dx, self.Dy, self.Dz)[a] > 1}
        if self.Dy > 1 and self.Dz > 1:
            nb |= {self._neighbour_yz(d, dy, dz) for dy in (-1, 1) for dz in (-1, 1)}
        return sorted(nb - {d})

    def _connect_peers(self):
        """table d -> {fi, rho, u, flags, sync} device pointers for every domain this process needs to read or signal"""
        peers = {d: dict(fi=dom.fi.device_ptr, rho=dom.rho.device_ptr, u=dom.u.device_ptr, flags=dom.flags.device_ptr, sync=dom.sync_array, xfer=dom.xfer, device=dom.device,
                         F=dom.F.device_ptr if dom.F is not None else None)
                 for d, dom in self.lbm_domain.items()}
        shared = ("fi", "rho", "u", "flags", "sync") + (("xfer",) if self.Dx > 1 else ()) + (("F",) if self.features & FORCE_FIELD else ())
        if self.comm is None:
            devs = sorted({dom.device for dom in self.lbm_domain.values()})
            for a in devs:
                for b in devs:
                    if a != b: self.lib.device_enable_peer(a, b)
        else:  # one process per GPU: exchange CUDA IPC handles of the five buffers
            (d, dom), = self.lbm_domain.items()
            mine = {}
            for k in shared:
                h = C.create_string_buffer(64)
                self.lib.ipc_get_handle(dom.device, peers[d][k], h)
                mine[k] = h.raw
            everyone = self.comm.allgather(mine)
            needed = set(self._sync_neighbours(d))
            for r in needed:
                entry = {"device": dom.device}
                for k in shared:
                    p = C.c_void_p()
                    self.lib.ipc_open_handle(dom.device, everyone[r][k], C.byref(p))
                    entry[k] = p.value
                peers[r] = entry
            self._ipc_opened = {r: peers[r] for r in needed}
        self._peers = peers
        # fused y/z halo delivery: per owned domain the 3x3 table of y/z neighbours' DDF buffers (include/fx3d.h: fx3d_stream_collide_fused)
        self._fused = {}
        if (self.Dy > 1 or self.Dz > 1) and not self.overlap and self.fuse_halo:
            for d, dom in self.lbm_domain.items():
                if self.lib.fused_halo_supported(C.byref(dom.lat)):
                    table = (C.c_void_p * 9)()
                    for dz in (-1, 0, 1):
                        for dy in (-1, 0, 1):
                            if (dy and self.Dy == 1) or (dz and self.Dz == 1): continue
                            table[(dy + 1) + 3 * (dz + 1)] = peers[self._neighbour_yz(d, dy, dz)]["fi"]
                    self._fused[d] = table

    # ---- halo exchange: replaces communicate_field, src/lbm.cpp:1355-1383 ----
    def _barrier(self, axis_or_all):
        """device-side rendezvous of every owned domain with its face neighbours (replaces the finish_queue barriers)"""
        self._seq += 1
        for d, dom in self.local_domains():
            nb = self._sync_neighbours(d)
            arr = (C.c_void_p * len(nb))(*[self._peers[r]["sync"] for r in nb])
            self.lib.rendezvous_signal(dom.device, arr, len(nb), d, self._seq, dom.stream)
        for d, dom in self.local_domains():
            nb = self._sync_neighbours(d)
            idx = (C.c_int * len(nb))(*nb)
            self.lib.rendezvous_wait(dom.device, dom.sync_array, idx, len(nb), self._seq, 20000, dom.stream)

    def _communicate(self, field, axes=(0, 1, 2)):
        for axis, Dn in enumerate((self.Dx, self.Dy, self.Dz)):
            if Dn <= 1 or axis not in axes: continue
            staged = axis == 0 and field == "fi"
            if staged:  # x faces: pack my two outgoing layers into my linear buffers first, so that the peer reads below are coalesced
                for d, dom in self.local_domains():
                    self.lib.transfer_extract_fi(C.byref(dom.lat), 0, dom.t, dom.xfer, dom.xfer + dom.xfer_bytes, dom.stream)
            self._barrier(axis)  # producers of this phase (stream_collide or the previous axis' pull) are done everywhere
            for d, dom in self.local_domains():
                p, m = self._peers[self._neighbour(d, axis, +1)], self._peers[self._neighbour(d, axis, -1)]
                if staged:  # my +x halo receives what the +x neighbour packed for its -x side, and vice versa
                    self.lib.transfer_insert_fi(C.byref(dom.lat), 0, dom.t, p["xfer"] + dom.xfer_bytes, m["xfer"], dom.stream)
                elif field == "fi":
                    self.lib.exchange_fi(C.byref(dom.lat), axis, dom.t, p["fi"], m["fi"], dom.stream)
                elif field == "flags":
                    self.lib.exchange_flags(C.byref(dom.lat), axis, p["flags"], m["flags"], dom.stream)
                elif field == "F":
                    self.lib.exchange_F(C.byref(dom.lat), axis, p["F"], m["F"], dom.stream)
                else:
                    self.lib.exchange_rho_u_flags(C.byref(dom.lat), axis, p["rho"], p["u"], p["flags"], m["rho"], m["u"], m["flags"], dom.stream)
        self._barrier(None)  # every neighbour has finished pulling from me before I overwrite what it read

    def communicate_fi(self): self._communicate("fi")  # src/lbm.cpp:1385-1387
    def communicate_rho_u_flags(self): self._communicate("rho_u_flags")  # src/lbm.cpp:1388-1390
    def communicate_flags(self): self._communicate("flags")  # src/lbm.cpp:1391-1393: one byte per face cell
    def communicate_F(self): self._communicate("F")  # src/lbm.cpp:1395-1397

    # ---- src/lbm.cpp:881-980 ----
    def initialize(self):
        if self.host_fields:
            if not self.benchmark: self._sanity_checks_initialization()
            for _, dom in self.local_domains(): dom.rho.enqueue_write_to_device(dom.stream)
            for _, dom in self.local_domains(): dom.u.enqueue_write_to_device(dom.stream)
            for _, dom in self.local_domains(): dom.flags.enqueue_write_to_device(dom.stream)
        if self.features & FORCE_FIELD:  # src/lbm.cpp:889-892
            for _, dom in self.local_domains(): dom.F.enqueue_write_to_device(dom.stream)
            if self.get_D() > 1: self.communicate_F()
        for _, dom in self.local_domains(): dom.increment_time_step()  # the communicate calls at initialization need an odd time step
        if self.get_D() > 1: self.communicate_rho_u_flags()
        for _, dom in self.local_domains(): dom.enqueue_initialize()
        if self.get_D() > 1:
            self.communicate_rho_u_flags()
            self.communicate_fi()
        for _, dom in self.local_domains(): dom.finish_queue()
        for _, dom in self.local_domains(): dom.reset_time_step()
        self.initialized = True

    def do_time_step(self):
        if self.overlap and self.get_D() > 1:
            # Shell first (the cells whose output the neighbours pull) on the main stream, followed there by the halo exchange;
            # the interior runs meanwhile on the second stream. The two regions never touch the same (cell, slot), the exchange
            # only reads what the shell wrote and only writes halo cells. Ordering between steps:
            # shell(n) after interior(n-1), interior(n) after shell(n) (hence after exchange(n-1), in stream order).
            devs = sorted(self._streams)
            for dev in devs: self.lib.stream_wait_event(dev, self._streams[dev], self._events[dev][1])
            for _, dom in self.local_domains(): dom.enqueue_stream_collide(REGION_SHELL)
            for dev in devs:
                self.lib.event_record(dev, self._events[dev][0], self._streams[dev])
                self.lib.stream_wait_event(dev, self._streams2[dev], self._events[dev][0])
            for _, dom in self.local_domains(): dom.enqueue_stream_collide(REGION_INTERIOR, stream=self._streams2[dom.device])
            for dev in devs: self.lib.event_record(dev, self._events[dev][1], self._streams2[dev])
        elif self._fused and len(self._fused) == len(self.lbm_domain):
            # y/z halo rows are delivered by the kernel itself (every stored row goes to the domain that reads it next); what remains is
            # one rendezvous -- nobody starts step t+1 before its neighbours have finished step t -- and, for x-decomposed grids, the x faces
            for d, dom in self.local_domains():
                self.lib.stream_collide_fused(C.byref(dom.lat), dom.t, dom.fx, dom.fy, dom.fz, self._fused[d], dom.stream)
            self._barrier(None)
            if self.Dx > 1: self._communicate("fi", axes=(0,))
            for _, dom in self.local_domains(): dom.increment_time_step()
            return
        else:
            for _, dom in self.local_domains(): dom.enqueue_stream_collide()
        if self.get_D() > 1: self.communicate_fi()
        for _, dom in self.local_domains(): dom.increment_time_step()

    def _join_streams(self):
        """everything enqueued after this on the main stream also follows the interior stream's work"""
        for dev in self._streams2: self.lib.stream_wait_event(dev, self._streams[dev], self._events[dev][1])

    def run(self, steps=max_ulong, total_steps=max_ulong, sync=True):
        """run(steps): first call initialises; run(0) initialises only (src/lbm.cpp:955-975). Unlike the reference there is no
        host synchronisation per step; with sync=True (default) run() returns after the device has finished."""
        if not self.initialized: self.initialize()
        if self.get_D() == 1 and steps > 0 and steps != max_ulong:
            (_, dom), = self.local_domains()
            self.lib.run_steps(C.byref(dom.lat), dom.t, steps, dom.fx, dom.fy, dom.fz, dom.stream)
            dom.increment_time_step(steps)
        else:
            i = 0
            while i < steps:
                self.do_time_step(); i += 1
            if self._streams2 and steps > 0: self._join_streams()
        if sync: self.finish()

    def finish(self):
        if self._streams2: self._join_streams()
        for _, dom in self.local_domains(): dom.finish_queue()
        if self.get_D() > 1:
            for _, dom in self.local_domains(): self.lib.rendezvous_check(dom.device, dom.sync_array, 64)

    def update_fields(self):
        if self._streams2: self._join_streams()
        for _, dom in self.local_domains(): dom.enqueue_update_fields()
        for _, dom in self.local_domains(): dom.finish_queue()

    def update_moving_boundaries(self):
        """mark / unmark the cells next to TYPE_S cells with non-zero velocity with TYPE_MS, after the boundary velocities in
        lbm.u / lbm.flags were changed and written to the device (src/lbm.cpp:1018-1027)"""
        if not (self.features & MOVING_BOUNDARIES): raise ValueError("update_moving_boundaries() needs the MOVING_BOUNDARIES extension")
        if self._streams2: self._join_streams()
        for _, dom in self.local_domains(): dom.enqueue_update_moving_boundaries()
        if self.get_D() > 1: self.communicate_flags()
        for _, dom in self.local_domains(): dom.finish_queue()

    # ---- FORCE_FIELD, src/lbm.cpp:986-1016 ----
    def _need_force_field(self):
        if not (self.features & FORCE_FIELD): raise ValueError("this call needs the FORCE_FIELD extension")
        if self._streams2: self._join_streams()

    def update_force_field(self):
        """forces of the fluid on the TYPE_S cells -> lbm.F on the device (read with lbm.F.read_from_device())"""
        self._need_force_field()
        for _, dom in self.local_domains(): dom.enqueue_update_force_field()
        for _, dom in self.local_domains(): dom.finish_queue()

    def _object_sums(self):
        for _, dom in self.local_domains(): dom.finish_queue()
        parts = [(np.array(dom.object_sum.host[:3], np.float32), int(dom.object_sum.host[3:4].view(np.uint32)[0])) for _, dom in self.local_domains()]
        if self.comm is not None:  # one process per GPU: the per-domain sums are added in domain order, as the reference does
            parts = [p for rank_parts in self.comm.allgather(parts) for p in rank_parts]
        total, cells = np.zeros(3, np.float32), 0
        for v, c in parts:
            total = (total + v).astype(np.float32); cells += c
        return total, cells

    def object_center_of_mass(self, flag_marker=TYPE_S):
        """centre of mass of all cells whose flag byte equals flag_marker (positions local to each domain, as in the reference)"""
        self._need_force_field()
        for _, dom in self.local_domains(): dom.enqueue_object_center_of_mass(flag_marker)
        total, cells = self._object_sums()
        return total / np.float32(cells)

    def object_force(self, flag_marker=TYPE_S):
        """total force of the fluid on all cells whose flag byte equals flag_marker"""
        self._need_force_field()
        for _, dom in self.local_domains(): dom.enqueue_object_force(flag_marker)
        return self._object_sums()[0]

    def object_torque(self, rotation_center, flag_marker=TYPE_S):
        self._need_force_field()
        for _, dom in self.local_domains(): dom.enqueue_object_torque(rotation_center, flag_marker)
        return self._object_sums()[0]

    # ---- GPU voxeliser, src/lbm.cpp:1074-1145 ----
    def voxelize_mesh_on_device(self, mesh, flag=TYPE_S, rotation_center=None, linear_velocity=(0.0, 0.0, 0.0), rotational_velocity=(0.0, 0.0, 0.0)):
        """mark the cells inside the closed triangle mesh with `flag` on every domain (kernel voxelize_mesh, src/kernel.cpp:2267-2345)"""
        from .mesh import voxelize_parameters
        bbu, direction = voxelize_parameters(mesh, mesh.center if rotation_center is None else rotation_center, linear_velocity, rotational_velocity)
        nbytes = mesh.triangle_number * 12
        for d, dom in self.local_domains():
            bufs = []
            for arr in (mesh.p0, mesh.p1, mesh.p2):
                p = C.c_void_p(); self.lib.malloc(dom.device, nbytes, C.byref(p))
                self.lib.memcpy_h2d(dom.device, p, arr.ctypes.data, nbytes, dom.stream, 1)
                bufs.append(p)
            self.lib.voxelize_mesh(C.byref(dom.lat), dom.Ox, dom.Oy, dom.Oz, direction, dom.t + 1, flag, bufs[0], bufs[1], bufs[2], bbu.ctypes.data, dom.stream)
            dom.finish_queue()
            for p in bufs: self.lib.free(dom.device, p)
        moving = any(v != 0.0 for v in linear_velocity) or any(v != 0.0 for v in rotational_velocity)
        if (self.features & MOVING_BOUNDARIES) and (flag & (TYPE_S | TYPE_E)) == TYPE_S and moving: self.update_moving_boundaries()
        if not self.initialized and self.host_fields:  # the host copies follow, so that initialize() does not overwrite the result
            self.flags.read_from_device(); self.u.read_from_device()

    def unvoxelize_mesh_on_device(self, mesh, flag=TYPE_S):
        """clear `flag` in the bounding box of the mesh (only needed when the box changes size between re-voxelisations), src/lbm.cpp:1090-1093"""
        for _, dom in self.local_domains():
            self.lib.unvoxelize_mesh(C.byref(dom.lat), dom.Ox, dom.Oy, dom.Oz, flag, *[C.c_float(float(v)) for v in (*mesh.pmin, *mesh.pmax)], dom.stream)
        for _, dom in self.local_domains(): dom.finish_queue()

    def voxelize_stl(self, path, center=None, rotation=None, size=0.0, flag=TYPE_S):
        """read a binary .stl file, fit it into the box (size 0), scale its longest side to `size` cells (size > 0) or by -size, and voxelise it
        (LBM::voxelize_stl, src/lbm.cpp:1130-1145; flags travel host -> device -> host around the kernel like there)"""
        from .mesh import read_stl
        mesh = read_stl(path, self.size(), self.center() if center is None else center, size, rotation)
        self.flags.write_to_device()
        self.voxelize_mesh_on_device(mesh, flag)
        self.flags.read_from_device()
        return mesh

    def reset(self): self.initialized = False

    # ---- getters, src/lbm.hpp:448-535 ----
    def get_Nx(self): return self.Nx
    def get_Ny(self): return self.Ny
    def get_Nz(self): return self.Nz
    def get_N(self): return self.Nx * self.Ny * self.Nz
    def get_Dx(self): return self.Dx
    def get_Dy(self): return self.Dy
    def get_Dz(self): return self.Dz
    def get_D(self): return self.Dx * self.Dy * self.Dz
    def _any(self): return self.local_domains()[0][1]
    def get_nu(self): return self._any().nu
    def get_tau(self): return 3.0 * self.get_nu() + 0.5
    def get_Re_max(self): return 0.57735027 * (self.Nx ** 2 + self.Ny ** 2 + self.Nz ** 2) ** 0.5 / self.get_nu()
    def get_t(self): return self._any().t
    def get_velocity_set(self): return self.velocity_set
    def get_fx(self): return self._any().fx
    def get_fy(self): return self._any().fy
    def get_fz(self): return self._any().fz

    def set_f(self, fx, fy, fz):
        for _, dom in self.local_domains(): dom.set_f(fx, fy, fz)

    def coordinates(self, n):
        t = n % (self.Nx * self.Ny)
        return t % self.Nx, t // self.Nx, n // (self.Nx * self.Ny)

    def index(self, x, y, z): return x + (y + z * self.Ny) * self.Nx
    def position(self, x, y, z): return (x - 0.5 * self.Nx + 0.5, y - 0.5 * self.Ny + 0.5, z - 0.5 * self.Nz + 0.5)
    def size(self): return (float(self.Nx), float(self.Ny), float(self.Nz))
    def center(self): return (0.5 * self.Nx - 0.5, 0.5 * self.Ny - 0.5, 0.5 * self.Nz - 0.5)
    def bandwidth_bytes_per_cell_device(self): return self.lib.bytes_per_cell_per_step(C.byref(self._any().lat))  # src/lbm.cpp:52

    def close(self):
        for _, dom in self.local_domains():
            try: dom.finish_queue()
            except Fx3dError: pass
        if self.comm is not None and getattr(self, "_ipc_opened", None):
            (_, dom), = self.lbm_domain.items()
            for entry in self._ipc_opened.values():
                for k in ("fi", "rho", "u", "flags", "sync", "xfer", "F"):
                    if entry.get(k): self.lib.ipc_close_handle(dom.device, entry[k])
            self._ipc_opened = None
            self.comm.barrier()
        for _, dom in self.local_domains(): dom.free()
        for dev, s in self._streams.items(): self.lib.stream_destroy(dev, s)
        for dev, s in self._streams2.items(): self.lib.stream_destroy(dev, s)
        for dev, (e1, e2) in self._events.items(): self.lib.event_destroy(dev, e1); self.lib.event_destroy(dev, e2)
        self.lbm_domain, self._streams, self._streams2, self._events = {}, {}, {}, {}


class TorchComm:
    """plumbing for one-process-per-GPU runs: object all-gather and barrier over torch.distributed (NCCL or gloo)"""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world_size = dist.get_rank(), dist.get_world_size()
        import os
        self.local_rank = int(os.environ.get("LOCAL_RANK", self.rank))

    def allgather(self, obj):
        out = [None] * self.world_size
        self.dist.all_gather_object(out, obj)
        return out

    def barrier(self): self.dist.barrier()
