"""Round-2 parity additions:
  * a14: the staged transfer_extract_rho_u_flags / transfer__insert_rho_u_flags pair (src/kernel.cpp:2133-2158) against the
    oracle's buffers, byte for byte, and the insert against the oracle's halo contents -- in emulation (CPU suite) and on the GPU
  * C3: D3Q27 TRT FP32, 64x128x64, sphere, six TYPE_E faces with u=(0,0.075,0), force (pattern of src/setup.cpp:614-618)
  * C1 parity at the BASELINE size: 256^3 FP32 perturbed initial condition, 20 steps, bit-exact against the oracle"""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from helpers import (ROOT, OracleBackend, HostSim, scenario, load_scenario, FP32, FP16S, FP16C, SRT, TRT, TYPE_S, TYPE_E)
from fluidx3d_b200 import capi
from fluidx3d_b200 import lbm as lbm_mod
from fluidx3d_b200.lbm import LBM

lbm_mod.VERBOSE = False
EMUL_SO = os.path.join(ROOT, "tests", "_build", "libfx3d_emul.so")


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "tests", "emul")], check=True, capture_output=True)
    return capi.Lib(EMUL_SO)


@pytest.fixture(scope="module")
def cuda():
    lib = capi.lib()
    assert lib.num_devices() >= 1
    return lib


def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def check_rho_u_flags_transfer(lib, Q, st, dims, D):
    sim = LBM(*dims, 0.05, Dx=D[0], Dy=D[1], Dz=D[2], velocity_set=Q, storage=st, lib=lib, devices=[0] * (D[0] * D[1] * D[2]))
    ref = HostSim(OracleBackend(Q, SRT, st, 0), *dims, *D, nu=0.05)
    scen = scenario(*dims, seed=12)
    sim.rho.set_global(scen[0]); [sim.u.set_global(scen[1][a], a) for a in range(3)]; sim.flags.set_global(scen[2])
    load_scenario(ref, *scen)
    sim.run(2); ref.run(2)
    sim.update_fields(); ref.update_fields()  # rho and u of every domain are current on the device / in the oracle's arrays
    d = D[0] * D[1] * D[2] - 1  # the last domain: non-zero offsets on every decomposed axis
    dom, rdom = sim.lbm_domain[d], ref.dom[d]
    nbytes = lib.transfer_bytes(C.byref(dom.lat))
    bp, bm = C.c_void_p(), C.c_void_p()
    lib.malloc(dom.device, nbytes, C.byref(bp)); lib.malloc(dom.device, nbytes, C.byref(bm))
    N = dom.get_N()
    for axis in range(3):
        if D[axis] == 1: continue
        A = [dom.Ny * dom.Nz, dom.Nz * dom.Nx, dom.Nx * dom.Ny][axis]
        # extract: 17 bytes per face cell, layout of extract_rho_u_flags (src/kernel.cpp:2133-2141)
        lib.transfer_extract_rho_u_flags(C.byref(dom.lat), axis, dom.t, bp, bm, dom.stream)
        hp, hm = np.zeros(nbytes, np.uint8), np.zeros(nbytes, np.uint8)
        lib.memcpy_d2h(dom.device, hp.ctypes.data, bp, nbytes, dom.stream, 1); lib.memcpy_d2h(dom.device, hm.ctypes.data, bm, nbytes, dom.stream, 1)
        ref.b.extract_ruf(axis, ref.t, rdom.buf_p, rdom.buf_m, rdom.rho, rdom.u, rdom.flags)
        assert np.array_equal(hp[:A * 17], rdom.buf_p[:A * 17]), f"axis {axis}: +buffer differs"
        assert np.array_equal(hm[:A * 17], rdom.buf_m[:A * 17]), f"axis {axis}: -buffer differs"
        # insert: feed both implementations the same synthetic buffers and compare the halo layers they write
        rng = np.random.default_rng(axis)
        sp, sm = rng.integers(0, 256, A * 17, dtype=np.uint8), rng.integers(0, 256, A * 17, dtype=np.uint8)
        for s in (sp, sm):  # keep the float payload finite: bit patterns of small floats
            s[:A * 16] = rng.uniform(-1, 1, A * 4).astype(np.float32).view(np.uint8)
        fullp, fullm = np.zeros(nbytes, np.uint8), np.zeros(nbytes, np.uint8)
        fullp[:A * 17], fullm[:A * 17] = sp, sm
        lib.memcpy_h2d(dom.device, bp, fullp.ctypes.data, nbytes, dom.stream, 1); lib.memcpy_h2d(dom.device, bm, fullm.ctypes.data, nbytes, dom.stream, 1)
        lib.transfer_insert_rho_u_flags(C.byref(dom.lat), axis, dom.t, bp, bm, dom.stream)
        rdom.buf_p[:A * 17], rdom.buf_m[:A * 17] = sp, sm
        ref.b.insert_ruf(axis, ref.t, rdom.buf_p, rdom.buf_m, rdom.rho, rdom.u, rdom.flags)
        g_rho, g_u, g_fl = np.zeros(N, np.float32), np.zeros(3 * N, np.float32), np.zeros(N, np.uint8)
        lib.memcpy_d2h(dom.device, g_rho.ctypes.data, dom.rho.device_ptr, 4 * N, dom.stream, 1)
        lib.memcpy_d2h(dom.device, g_u.ctypes.data, dom.u.device_ptr, 12 * N, dom.stream, 1)
        lib.memcpy_d2h(dom.device, g_fl.ctypes.data, dom.flags.device_ptr, N, dom.stream, 1)
        assert np.array_equal(bits(g_rho), bits(rdom.rho)) and np.array_equal(bits(g_u), bits(rdom.u)) and np.array_equal(g_fl, rdom.flags), f"axis {axis}: inserted halo differs"
    lib.free(dom.device, bp); lib.free(dom.device, bm)
    sim.close()


@pytest.mark.parametrize("Q,st,dims,D", [(19, FP32, (12, 8, 6), (2, 2, 2)), (27, FP16C, (10, 6, 8), (2, 1, 2))], ids=["q19-fp32-d222", "q27-fp16c-d212"])
def test_transfer_rho_u_flags_matches_oracle_emulated(emul, Q, st, dims, D):
    check_rho_u_flags_transfer(emul, Q, st, dims, D)


@pytest.mark.gpu
@pytest.mark.parametrize("Q,st,dims,D", [(19, FP32, (12, 8, 6), (2, 2, 2)), (27, FP16C, (10, 6, 8), (2, 1, 2)), (19, FP16S, (64, 48, 32), (2, 2, 2))], ids=["q19-fp32-d222", "q27-fp16c-d212", "q19-fp16s-64x48x32-d222"])
def test_transfer_rho_u_flags_matches_oracle_gpu(cuda, Q, st, dims, D):
    check_rho_u_flags_transfer(cuda, Q, st, dims, D)


def windtunnel_scene(Nx, Ny, Nz):
    """sphere of radius Nx/8 at (Nx/2, Ny/4, Nz/2), all six faces TYPE_E with u=(0,0.075,0), fluid starts at the same velocity"""
    zz, yy, xx = np.meshgrid(np.arange(Nz), np.arange(Ny), np.arange(Nx), indexing="ij", sparse=True)
    flags = np.zeros((Nz, Ny, Nx), np.uint8)
    flags[(xx - Nx / 2) ** 2 + (yy - Ny / 4) ** 2 + (zz - Nz / 2) ** 2 <= (Nx / 8) ** 2] = TYPE_S
    for sl in [np.s_[0, :, :], np.s_[-1, :, :], np.s_[:, 0, :], np.s_[:, -1, :], np.s_[:, :, 0], np.s_[:, :, -1]]:
        flags[sl] = TYPE_E
    rho = np.ones((Nz, Ny, Nx), np.float32)
    zero = np.zeros((Nz, Ny, Nx), np.float32)
    uy = np.where(flags == TYPE_S, 0.0, 0.075).astype(np.float32)
    return rho, [zero, uy, zero.copy()], flags


def check_windtunnel(lib, dims, D, steps, storage=FP32):
    Nx, Ny, Nz = dims
    nu = Nx * 0.075 / 1e4  # Re = 10^4 on Nx (units.nu_from_Re)
    f = (0.0, 1e-6, 0.0)
    scen = windtunnel_scene(Nx, Ny, Nz)
    sim = LBM(Nx, Ny, Nz, nu, *f, Dx=D[0], Dy=D[1], Dz=D[2], velocity_set=27, collision=TRT, storage=storage, features=3, lib=lib, devices=[0] * (D[0] * D[1] * D[2]))
    sim.rho.set_global(scen[0]); [sim.u.set_global(scen[1][a], a) for a in range(3)]; sim.flags.set_global(scen[2])
    sim.run(steps)
    for m in (sim.rho, sim.u, sim.flags): m.read_from_device()
    got = (sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2), sim.flags.get_global())
    sim.close()
    ref = HostSim(OracleBackend(27, TRT, storage, 3), Nx, Ny, Nz, *D, nu=nu, fx=f[0], fy=f[1], fz=f[2])
    load_scenario(ref, *scen)
    ref.run(steps)
    for name, a, b in zip(("rho", "ux", "uy", "uz", "flags"), got, ref.fields()):
        assert np.array_equal(bits(a), bits(b)), f"{name} differs from the oracle in {int(np.sum(bits(a) != bits(b)))} cells"
    assert float(np.max(np.abs(got[1]))) > 1e-4, "the sphere did not deflect the flow"


def test_windtunnel_fixture_emulated(emul):
    check_windtunnel(emul, (16, 32, 16), (1, 1, 1), 4)


@pytest.mark.gpu
@pytest.mark.parametrize("D", [(1, 1, 1), (2, 2, 1)], ids=["d111", "d221"])
def test_windtunnel_fixture_c3_gpu(cuda, D):
    check_windtunnel(cuda, (64, 128, 64), D, 30)


@pytest.mark.gpu
@pytest.mark.parametrize("storage", [FP32, FP16S], ids=["fp32", "fp16s"])
def test_perturbed_256_cubed_bit_exact_gpu(cuda, storage):
    """BASELINE configs[0] size: 256^3, perturbed rho/u (no solids: the benchmark box), 20 steps, bit-exact against the oracle"""
    N, steps = 256, 20
    scen = scenario(N, N, N, seed=1, solid_frac=0.0)
    sim = LBM(N, N, N, 1.0, velocity_set=19, collision=SRT, storage=storage, lib=cuda)
    sim.rho.set_global(scen[0]); [sim.u.set_global(scen[1][a], a) for a in range(3)]; sim.flags.set_global(scen[2])
    sim.run(steps)
    for m in (sim.rho, sim.u): m.read_from_device()
    got = (sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2))
    sim.close()
    ref = HostSim(OracleBackend(19, SRT, storage, 0), N, N, N, nu=1.0)
    load_scenario(ref, *scen)
    ref.run(steps)
    for name, a, b in zip(("rho", "ux", "uy", "uz"), got, ref.fields()):
        assert np.array_equal(bits(a), bits(b)), f"{name} differs from the oracle in {int(np.sum(bits(a) != bits(b)))} cells"
