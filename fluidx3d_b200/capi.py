"""ctypes binding of libfx3d_cuda.so (include/fx3d.h). The library is CUDA-only: there is no CPU path, and loading
fails loudly when the extension has not been built (run `python -c "import __graft_entry__ as g; g.build()"`)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.environ.get("FX3D_LIB", os.path.join(HERE, "libfx3d_cuda.so"))  # FX3D_LIB: alternative build of the same CUDA library (tuning experiments)

FP32, FP16S, FP16C = 0, 1, 2
SRT, TRT = 0, 1
VOLUME_FORCE, EQUILIBRIUM_BOUNDARIES, UPDATE_FIELDS, SUBGRID, MOVING_BOUNDARIES, FORCE_FIELD = 1, 2, 4, 8, 16, 32
REGION_ALL, REGION_SHELL, REGION_INTERIOR = 0, 1, 2
TYPE_S, TYPE_E = 0x01, 0x02  # src/defines.hpp:52-53
OK, ERR_NO_DEVICE, ERR_INVALID, ERR_OUT_OF_MEMORY, ERR_CUDA, ERR_TIMEOUT = 0, -1, -2, -3, -4, -5


class Fx3dError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fx3d error {code}: {msg}")
        self.code = code


class DeviceInfo(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("id", C.c_int), ("cc_major", C.c_int), ("cc_minor", C.c_int), ("sm_count", C.c_int),
                ("clock_mhz", C.c_int), ("memory_bytes", C.c_uint64), ("l2_bytes", C.c_uint64), ("tflops_fp32", C.c_float)]


class Lattice(C.Structure):
    _fields_ = [("device", C.c_int), ("Nx", C.c_uint32), ("Ny", C.c_uint32), ("Nz", C.c_uint32),
                ("Dx", C.c_uint32), ("Dy", C.c_uint32), ("Dz", C.c_uint32),
                ("velocity_set", C.c_uint32), ("collision", C.c_uint32), ("storage", C.c_uint32), ("features", C.c_uint32),
                ("w", C.c_float), ("fi", C.c_void_p), ("rho", C.c_void_p), ("u", C.c_void_p), ("flags", C.c_void_p), ("F", C.c_void_p)]


_VP, _U64, _U32, _F, _I, _SZ = C.c_void_p, C.c_uint64, C.c_uint32, C.c_float, C.c_int, C.c_size_t
_LP = C.POINTER(Lattice)
# name -> (restype, argtypes, returns_status)
_SIGS = {
    "fx3d_last_error": (C.c_char_p, [], False),
    "fx3d_device_count": (_I, [C.POINTER(_I)], True),
    "fx3d_device_get_info": (_I, [_I, C.POINTER(DeviceInfo)], True),
    "fx3d_device_enable_peer": (_I, [_I, _I], True),
    "fx3d_device_sync": (_I, [_I], True),
    "fx3d_stream_create": (_I, [_I, C.POINTER(_VP)], True),
    "fx3d_stream_destroy": (_I, [_I, _VP], True),
    "fx3d_stream_sync": (_I, [_I, _VP], True),
    "fx3d_event_create": (_I, [_I, C.POINTER(_VP)], True),
    "fx3d_event_destroy": (_I, [_I, _VP], True),
    "fx3d_event_record": (_I, [_I, _VP, _VP], True),
    "fx3d_event_sync": (_I, [_I, _VP], True),
    "fx3d_event_elapsed_ms": (_I, [_VP, _VP, C.POINTER(_F)], True),
    "fx3d_stream_wait_event": (_I, [_I, _VP, _VP], True),
    "fx3d_malloc": (_I, [_I, _SZ, C.POINTER(_VP)], True),
    "fx3d_free": (_I, [_I, _VP], True),
    "fx3d_host_alloc": (_I, [_SZ, C.POINTER(_VP)], True),
    "fx3d_host_free": (_I, [_VP], True),
    "fx3d_memcpy_h2d": (_I, [_I, _VP, _VP, _SZ, _VP, _I], True),
    "fx3d_memcpy_d2h": (_I, [_I, _VP, _VP, _SZ, _VP, _I], True),
    "fx3d_memset": (_I, [_I, _VP, _I, _SZ, _VP], True),
    "fx3d_fill_f32": (_I, [_I, _VP, _F, _SZ, _VP], True),
    "fx3d_ipc_get_handle": (_I, [_I, _VP, _VP], True),
    "fx3d_ipc_open_handle": (_I, [_I, _VP, C.POINTER(_VP)], True),
    "fx3d_ipc_close_handle": (_I, [_I, _VP], True),
    "fx3d_fi_bytes": (_SZ, [_LP], False),
    "fx3d_relaxation_rate": (_F, [_F], False),
    "fx3d_bytes_per_cell_per_step": (_U32, [_LP], False),
    "fx3d_initialize": (_I, [_LP, _VP], True),
    "fx3d_update_moving_boundaries": (_I, [_LP, _VP], True),
    "fx3d_stream_collide_launches": (_I, [_I, C.POINTER(_U64)], True),
    "fx3d_set_interior_reserve": (_I, [_I], True),
    "fx3d_stream_collide": (_I, [_LP, _U64, _F, _F, _F, _I, _VP], True),
    "fx3d_stream_collide_fused": (_I, [_LP, _U64, _F, _F, _F, C.POINTER(_VP), _VP], True),
    "fx3d_fused_halo_supported": (_I, [_LP], False),
    "fx3d_update_fields": (_I, [_LP, _U64, _F, _F, _F, _VP], True),
    "fx3d_update_force_field": (_I, [_LP, _U64, _VP], True),
    "fx3d_reset_force_field": (_I, [_LP, _VP], True),
    "fx3d_object_scratch_bytes": (_SZ, [_LP], False),
    "fx3d_object_center_of_mass": (_I, [_LP, C.c_uint8, _VP, _VP, _VP], True),
    "fx3d_object_force": (_I, [_LP, C.c_uint8, _VP, _VP, _VP], True),
    "fx3d_object_torque": (_I, [_LP, C.c_uint8, _F, _F, _F, _VP, _VP, _VP], True),
    "fx3d_voxelize_mesh": (_I, [_LP, _I, _I, _I, _U32, _U64, C.c_uint8, _VP, _VP, _VP, _VP, _VP], True),
    "fx3d_unvoxelize_mesh": (_I, [_LP, _I, _I, _I, C.c_uint8, _F, _F, _F, _F, _F, _F, _VP], True),
    "fx3d_run_steps": (_I, [_LP, _U64, _U64, _F, _F, _F, _VP], True),
    "fx3d_set_kernel_variant": (_I, [_I], True),
    "fx3d_launch_count": (_I, [C.POINTER(_U64)], True),
    "fx3d_transfer_bytes": (_SZ, [_LP], False),
    "fx3d_transfer_extract_fi": (_I, [_LP, _U32, _U64, _VP, _VP, _VP], True),
    "fx3d_transfer_insert_fi": (_I, [_LP, _U32, _U64, _VP, _VP, _VP], True),
    "fx3d_transfer_extract_rho_u_flags": (_I, [_LP, _U32, _U64, _VP, _VP, _VP], True),
    "fx3d_transfer_insert_rho_u_flags": (_I, [_LP, _U32, _U64, _VP, _VP, _VP], True),
    "fx3d_transfer_extract_flags": (_I, [_LP, _U32, _U64, _VP, _VP, _VP], True),
    "fx3d_transfer_insert_flags": (_I, [_LP, _U32, _U64, _VP, _VP, _VP], True),
    "fx3d_transfer_extract_F": (_I, [_LP, _U32, _U64, _VP, _VP, _VP], True),
    "fx3d_transfer_insert_F": (_I, [_LP, _U32, _U64, _VP, _VP, _VP], True),
    "fx3d_exchange_flags": (_I, [_LP, _U32, _VP, _VP, _VP], True),
    "fx3d_exchange_F": (_I, [_LP, _U32, _VP, _VP, _VP], True),
    "fx3d_exchange_fi": (_I, [_LP, _U32, _U64, _VP, _VP, _VP], True),
    "fx3d_exchange_rho_u_flags": (_I, [_LP, _U32, _VP, _VP, _VP, _VP, _VP, _VP, _VP], True),
    "fx3d_rendezvous_signal": (_I, [_I, C.POINTER(_VP), _I, _I, _U64, _VP], True),
    "fx3d_rendezvous_wait": (_I, [_I, _VP, C.POINTER(_I), _I, _U64, _I, _VP], True),
    "fx3d_rendezvous_check": (_I, [_I, _VP, _I], True),
    "fx3d_codec_encode": (_I, [_I, _I, _VP, _VP, _SZ, _VP], True),
    "fx3d_codec_decode": (_I, [_I, _I, _VP, _VP, _SZ, _VP], True),
    "fx3d_codec_fp16c_exhaustive": (_I, [_I, C.POINTER(_U64), C.POINTER(_U32)], True),
    "fx3d_selftest_division": (_I, [_I, _U64, C.POINTER(_U64)], True),
    "fx3d_selftest_packed_math": (_I, [_I, _U64, C.POINTER(_U64)], True),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


class Lib:
    """Thin checked wrapper: every status-returning entry point raises Fx3dError on failure."""

    def __init__(self, path=None, require=EXPORTED_SYMBOLS):
        self.path = path or DEFAULT_LIB
        if not os.path.exists(self.path):
            raise Fx3dError(ERR_NO_DEVICE, f"{self.path} is not built; the CUDA extension is required (no CPU fallback). "
                                           "Build it with __graft_entry__.build() or `make -C fluidx3d_b200/csrc`.")
        self.cdll = C.CDLL(self.path)
        for name, (res, args, status) in _SIGS.items():
            if not hasattr(self.cdll, name):
                if name in require:
                    raise Fx3dError(ERR_INVALID, f"{self.path} does not export {name}")
                continue
            fn = getattr(self.cdll, name)
            fn.restype, fn.argtypes = res, args
            setattr(self, name[5:], self._checked(fn) if status else fn)

    def _checked(self, fn):
        def call(*a):
            rc = fn(*a)
            if rc != 0:
                raise Fx3dError(rc, self.cdll.fx3d_last_error().decode(errors="replace"))
            return rc
        return call

    # conveniences
    def num_devices(self):
        n = C.c_int(0)
        self.device_count(C.byref(n))
        return n.value

    def info(self, device):
        out = DeviceInfo()
        self.device_get_info(device, C.byref(out))
        return out

    def launches(self):
        v = C.c_uint64(0)
        self.launch_count(C.byref(v))
        return v.value

    KERNEL_KINDS = ("k_stream_collide_v1", "k_stream_collide_vec", "k_stream_collide_pipe", "k_stream_collide_tma", "k_stream_collide_tma_seg", "k_stream_collide_hyb", "k_stream_collide_occ", "k_stream_collide_row")

    def kernel_kind_counts(self):
        """stream_collide launches so far per kernel kind (same order as KERNEL_KINDS)"""
        out = []
        for k in range(len(self.KERNEL_KINDS)):
            v = C.c_uint64(0)
            self.stream_collide_launches(k, C.byref(v))
            out.append(v.value)
        return out


_default = None


def lib():
    """the product library (built in-tree by __graft_entry__.build())"""
    global _default
    if _default is None:
        _default = Lib()
    return _default
