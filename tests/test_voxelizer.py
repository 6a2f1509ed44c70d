"""GPU voxeliser (kernel voxelize_mesh / unvoxelize_mesh src/kernel.cpp:2267-2357, read_stl src/utilities.hpp:4530-4581,
LBM::voxelize_stl / voxelize_mesh_on_device src/lbm.cpp:275-327,1074-1145).
  * the oracle's restatement equals the reference's own kernel source compiled natively (oracle/_ref), flags AND velocities, for single and
    multi-domain grids, every ray direction, resting and moving bodies
  * the CUDA path (emulated on CPU, real on the GPU) equals the oracle bit for bit; multi-domain flags equal single-domain flags
    (README.md:154 claims binary identity)
  * the product's read_stl equals the test-side restatement; a committed golden flag field pins all of it"""
import os
import subprocess
import numpy as np
import pytest
import helpers as H
from helpers import ROOT, OracleBackend, RefBackend, HostSim, FP32, FP16S, FP16C, SRT, TRT, TYPE_S, ref_available
from fluidx3d_b200 import capi
from fluidx3d_b200 import lbm as lbm_mod
from fluidx3d_b200 import mesh as mesh_mod
from fluidx3d_b200.lbm import LBM

lbm_mod.VERBOSE = False
STL = os.path.join(ROOT, "tests", "golden", "torus.stl")
GOLDEN = os.path.join(ROOT, "tests", "golden", "voxelize_torus_40x36x28.npz")
EMUL_SO = os.path.join(ROOT, "tests", "_build", "libfx3d_emul.so")
ROT = [[1, 0, 0], [0, 0.8, -0.6], [0, 0.6, 0.8]]
DIMS = (40, 36, 28)


def torus(dims, size_frac=0.8, rotation=ROT):
    Nx, Ny, Nz = dims
    return H.read_stl(STL, (Nx, Ny, Nz), (0.5 * Nx - 0.5, 0.5 * Ny - 0.5, 0.5 * Nz - 0.5), size_frac * Nx, rotation=rotation)


def oracle_voxelize(backend, dims, D, mesh, **kw):
    sim = HostSim(backend, *dims, *D, nu=0.05)
    sim.voxelize_mesh(mesh, **kw)
    return sim.get_global("flags"), [sim.get_global("u", a) for a in range(3)]


@pytest.mark.skipif(not ref_available(19, SRT, FP32, 0), reason="oracle/_ref not built")
@pytest.mark.parametrize("D", [(1, 1, 1), (2, 1, 2), (2, 2, 2)], ids=lambda d: "d%d%d%d" % d)
@pytest.mark.parametrize("kw", [{}, {"linear_velocity": (0.01, 0.0, -0.02)}, {"rotational_velocity": (0.0, 0.002, 0.001)}], ids=["rest", "moving", "rotating"])
def test_oracle_voxelizer_equals_reference_device_code(D, kw):
    mesh = torus(DIMS)
    f_o, u_o = oracle_voxelize(OracleBackend(19, SRT, FP32, 0), DIMS, D, mesh, **kw)
    f_r, u_r = oracle_voxelize(RefBackend(19, SRT, FP32, 0), DIMS, D, mesh, **kw)
    assert np.array_equal(f_o, f_r) and int((f_o == TYPE_S).sum()) > 1000
    for a, b in zip(u_o, u_r): assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_oracle_voxelizer_golden_and_domain_independence():
    mesh = torus(DIMS)
    single, _ = oracle_voxelize(OracleBackend(19, SRT, FP32, 0), DIMS, (1, 1, 1), mesh)
    for D in [(2, 1, 1), (1, 2, 2), (2, 2, 2)]:
        multi, _ = oracle_voxelize(OracleBackend(19, SRT, FP32, 0), DIMS, D, mesh)
        assert np.array_equal(single, multi), f"multi-domain voxelisation {D} differs from the single-domain one"
    assert np.array_equal(single, np.load(GOLDEN)["flags"])  # made from oracle/_ref by oracle/make_golden.py


def test_product_read_stl_matches_restatement():
    for size, rot in ((0.0, None), (25.0, ROT), (-3.5, ROT)):
        a = H.read_stl(STL, DIMS, (19.5, 17.5, 13.5), size, rotation=rot)
        b = mesh_mod.read_stl(STL, DIMS, (19.5, 17.5, 13.5), size, rotation=rot)
        for x, y in ((a.p0, b.p0), (a.p1, b.p1), (a.p2, b.p2), (a.pmin, b.pmin), (a.pmax, b.pmax)):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
        pa, da = H.voxelize_parameters(a, a.center, (0, 0, 0), (0, 0, 0)); pb, db = mesh_mod.voxelize_parameters(b, b.center, (0, 0, 0), (0, 0, 0))
        assert da == db and np.array_equal(pa.view(np.uint32), pb.view(np.uint32))


def product_voxelize(lib, v, dims, D, steps_before=0, **kw):
    """voxelise through the product's host API; with steps_before > 0 a moving body is re-voxelised after the fluid has run (cells it leaves restart from equilibrium)"""
    Q, coll, st, feat = v
    sim = LBM(*dims, 0.05, Dx=D[0], Dy=D[1], Dz=D[2], velocity_set=Q, collision=coll, storage=st, features=feat, lib=lib, devices=[0] * (D[0] * D[1] * D[2]))
    ref = HostSim(OracleBackend(Q, coll, st, feat), *dims, *D, nu=0.05)
    rho, u, flags = H.scenario(*dims, seed=5, solid_frac=0.0)
    sim.rho.set_global(rho); [sim.u.set_global(u[a], a) for a in range(3)]
    H.load_scenario(ref, rho, u, flags)
    m1 = mesh_mod.read_stl(STL, dims, sim.center(), 0.7 * dims[0], ROT)
    sim.rho.write_to_device(); sim.flags.write_to_device(); sim.u.write_to_device()
    sim.voxelize_mesh_on_device(m1, **kw)
    ref.voxelize_mesh(H.Mesh(m1.p0, m1.p1, m1.p2, m1.center), **kw)
    if steps_before:
        sim.run(steps_before); ref.run(steps_before)
        m1.translate((1.0, 0.0, -1.0))  # the body has moved: cells it left become fluid again
        sim.voxelize_mesh_on_device(m1, **kw)
        ref.voxelize_mesh(H.Mesh(m1.p0, m1.p1, m1.p2, m1.center), **kw)
        sim.run(2); ref.run(2)
    for m in (sim.rho, sim.u, sim.flags): m.read_from_device()
    got = (sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2), sim.flags.get_global())
    sim.close()
    want = ref.fields() if steps_before else (ref.get_global("rho"), ref.get_global("u", 0), ref.get_global("u", 1), ref.get_global("u", 2), ref.get_global("flags"))
    assert int((want[4] & 3 == TYPE_S).sum()) > 500
    for name, a, b in zip(("rho", "ux", "uy", "uz", "flags"), got, want):
        a = a.view(np.uint32) if a.dtype == np.float32 else a; b = b.view(np.uint32) if b.dtype == np.float32 else b
        assert np.array_equal(a, b), f"{name} differs in {int(np.sum(a != b))} cells"


CASES = [((19, SRT, FP32, 0), (1, 1, 1), 0, {}), ((19, SRT, FP16S, 0), (2, 1, 2), 0, {}), ((27, TRT, FP16C, 0), (2, 2, 2), 0, {}),
         ((19, SRT, FP32, 16), (1, 1, 1), 3, {"linear_velocity": (0.02, 0.0, -0.02)}), ((19, SRT, FP16S, 16), (1, 2, 2), 3, {"linear_velocity": (0.02, 0.0, -0.02)}),
         ((19, SRT, FP32, 16), (1, 1, 1), 0, {"rotational_velocity": (0.0, 0.002, 0.001)})]
IDS = [f"q{c[0][0]}s{c[0][2]}f{c[0][3]}-d{''.join(map(str, c[1]))}-{'re' if c[2] else ''}{'-'.join(c[3]) or 'rest'}" for c in CASES]


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "tests", "emul")], check=True, capture_output=True)
    return capi.Lib(EMUL_SO)


@pytest.mark.parametrize("v,D,steps,kw", CASES, ids=IDS)
def test_voxelizer_matches_oracle_emulated(emul, v, D, steps, kw):
    product_voxelize(emul, v, (32, 24, 20), D, steps, **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("v,D,steps,kw", CASES, ids=IDS)
def test_voxelizer_matches_oracle_gpu(v, D, steps, kw):
    product_voxelize(capi.lib(), v, (96, 64, 48), D, steps, **kw)


@pytest.mark.gpu
def test_voxelize_stl_through_host_api_gpu():
    """LBM.voxelize_stl: read_stl + write flags + kernel + read flags; 128^3, the wind-tunnel body placement of SURVEY 8d C3"""
    sim = LBM(128, 256, 128, 0.01, velocity_set=19, storage=FP16S, lib=capi.lib())
    mesh = sim.voxelize_stl(STL, center=(63.5, 63.5, 63.5), rotation=ROT, size=48.0)
    got = sim.flags.get_global()
    ref = HostSim(OracleBackend(19, SRT, FP16S, 0), 128, 256, 128, nu=0.01)
    ref.voxelize_mesh(H.Mesh(mesh.p0, mesh.p1, mesh.p2, mesh.center))
    sim.close()
    assert np.array_equal(got, ref.get_global("flags")) and int((got == TYPE_S).sum()) > 5000
