"""The C++ host surface (fluidx3d_b200/host: LBM, LBM_Domain, Memory<T>, Memory_Container) EXECUTED, not just compiled: the
scene tests/scenes/file_scene.cpp -- the same source that drives the untouched reference program in oracle/_ref/opencl -- is
built against the host classes, run, and its dumped fields are compared with the oracle bit for bit.

  * CPU suite: linked against the test-only emulation of the C ABI (tests/_build/libfx3d_emul.so): product kernels as OS threads
  * -m gpu:    linked against the product library libfx3d_cuda.so on the B200

Covers FP32 / FP16S / FP16C, domain decompositions on one device, TYPE_E + VOLUME_FORCE, and MOVING_BOUNDARIES with a mid-run
update_moving_boundaries() (the reference's lid-driven cavity mechanism, src/lbm.cpp:1018-1027)."""
import os
import subprocess
import numpy as np
import pytest
from helpers import (ROOT, OracleBackend, HostSim, scenario, load_scenario, FP32, FP16S, FP16C, SRT, TRT)

HOST = os.path.join(ROOT, "fluidx3d_b200", "host")
BUILD = os.path.join(ROOT, "tests", "_build")
SCENE = os.path.join(ROOT, "tests", "scenes", "file_scene.cpp")
EMUL_SO = os.path.join(BUILD, "libfx3d_emul.so")
PRODUCT_SO = os.path.join(ROOT, "fluidx3d_b200", "libfx3d_cuda.so")
ST_DEF = {FP32: [], FP16S: ["-DFP16S"], FP16C: ["-DFP16C"]}
FEAT_DEF = {1: "-DVOLUME_FORCE", 2: "-DEQUILIBRIUM_BOUNDARIES", 4: "-DUPDATE_FIELDS", 8: "-DSUBGRID", 16: "-DMOVING_BOUNDARIES", 32: "-DFORCE_FIELD"}
STL = os.path.join(ROOT, "tests", "golden", "torus.stl")
ROT = [[1, 0, 0], [0, 0.8, -0.6], [0, 0.6, 0.8]]
TYPE_SX = 0x41  # TYPE_S|TYPE_X

# (Q, collision, storage, features), grid, domains, steps, force, moving-boundary update "z,uy" or None
CASES = [((19, SRT, FP32, 0), (16, 8, 6), (1, 1, 1), 4, None, None),
         ((19, SRT, FP16S, 0), (32, 16, 4), (1, 1, 1), 5, None, None),       # whole-row tiles: bulk-copy kernel
         ((19, SRT, FP16C, 0), (24, 6, 6), (2, 1, 1), 4, None, None),        # x decomposition on one device
         ((27, TRT, FP32, 3), (16, 8, 6), (1, 2, 2), 4, (1e-4, -2e-4, 3e-4), None),
         ((19, SRT, FP32, 16), (16, 8, 6), (1, 1, 2), 6, None, "1,0.05"),    # MOVING_BOUNDARIES with update_moving_boundaries() after 3 steps
         ((19, SRT, FP16S, 8), (16, 8, 6), (1, 1, 1), 4, None, None),        # SUBGRID
         ((19, SRT, FP16S, 0), (64, 32, 8), (2, 2, 2), 5, None, None),       # whole-row tiles on every domain: y/z halo delivery fused into the kernel, x faces exchanged
         ((19, TRT, FP32, 3), (32, 32, 8), (1, 2, 2), 4, (1e-4, -2e-4, 3e-4), None),
         # voxelize_stl (read_stl + GPU voxeliser) and FORCE_FIELD: a torus in a box, boundary forces, object_force / object_center_of_mass / object_torque
         ((19, SRT, FP32, 34), (24, 20, 16), (1, 1, 1), 6, None, "stl"), ((19, SRT, FP16S, 32), (32, 24, 16), (2, 1, 2), 5, None, "stl")]
GPU_CASES = CASES + [((19, SRT, FP16S, 0), (512, 8, 8), (1, 1, 1), 6, None, None), ((19, SRT, FP32, 0), (128, 64, 32), (2, 2, 2), 10, None, None),
                     ((19, SRT, FP32, 16), (64, 64, 64), (1, 1, 1), 20, None, "63,0.1")]  # lid-driven cavity mechanism at a realistic size


def case_id(c):
    (Q, coll, st, feat), dims, D, steps, f, mb = c
    return f"q{Q}c{coll}s{st}f{feat}-{'x'.join(map(str, dims))}-d{''.join(map(str, D))}"


def build_scene(v, lib_so, tag):
    Q, coll, st, feat = v
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, f"host_scene_{tag}_q{Q}c{coll}s{st}f{feat}")
    defs = ["-DFX3D_CUSTOM_DEFINES", f"-DD3Q{Q}", "-DSRT" if coll == SRT else "-DTRT"] + ST_DEF[st] + [d for b, d in FEAT_DEF.items() if feat & b]
    src = [os.path.join(HOST, f) for f in ("main.cpp", "lbm.cpp", "info.cpp", "shapes.cpp")] + [SCENE]
    newest = max(os.path.getmtime(p) for p in src + [os.path.join(HOST, h) for h in os.listdir(HOST) if h.endswith(".hpp")] + [lib_so])
    if not os.path.exists(exe) or os.path.getmtime(exe) < newest:
        cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-pthread", "-w", "-I" + HOST] + defs + ["-o", exe] + src + [lib_so, "-Wl,-rpath," + os.path.dirname(lib_so)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    return exe


def run_case(case, lib_so, tag, tmp_path):
    v, dims, D, steps, f, mb = case
    Q, coll, st, feat = v
    exe = build_scene(v, lib_so, tag)
    Nx, Ny, Nz = dims
    rho, u, flags = scenario(Nx, Ny, Nz, seed=9, eq_frac=0.03 if feat & 2 else 0.0)
    stl = mb == "stl"
    if stl: mb = None
    if mb:  # a solid plane whose velocity will be switched on mid-run
        z = int(mb.split(",")[0]); flags[z, :, :] = 1
        for a in range(3): u[a][z, :, :] = 0.0
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as fh:
        np.array([Nx, Ny, Nz], np.uint32).tofile(fh); rho.tofile(fh); [u[a].tofile(fh) for a in range(3)]; flags.tofile(fh)
    nu = 0.05
    env = dict(os.environ, FX3D_REF_IN=fin, FX3D_REF_OUT=fout, FX3D_REF_STEPS=str(steps), FX3D_REF_NU=repr(nu), FX3D_REF_D="%d,%d,%d" % D)
    if f: env["FX3D_REF_F"] = ",".join(repr(x) for x in f)
    if mb: env["FX3D_REF_MB"] = mb
    stl_size = 0.7 * Nx
    if stl: env["FX3D_REF_STL"] = f"{STL},{stl_size!r}"; env["FX3D_REF_FF"] = "1"
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and os.path.exists(fout), (r.stdout[-1500:], r.stderr[-1500:])
    N = Nx * Ny * Nz
    raw = np.fromfile(fout, np.uint8)
    got = list(raw[:16 * N].view(np.float32).reshape(4, Nz, Ny, Nx)) + [raw[16 * N:17 * N].reshape(Nz, Ny, Nx)]
    got_ff = raw[17 * N:].view(np.float32) if stl else None
    # the oracle, driven through the same sequence
    ref = HostSim(OracleBackend(Q, coll, st, feat), Nx, Ny, Nz, *D, nu=nu, fx=(f or (0, 0, 0))[0], fy=(f or (0, 0, 0))[1], fz=(f or (0, 0, 0))[2])
    if stl:  # LBM::voxelize_stl on the empty box (mesh at the box centre, rotated, longest side = size cells); the input fields go to the cells outside the body
        import helpers as H
        ref.voxelize_mesh(H.read_stl(STL, (Nx, Ny, Nz), (0.5 * Nx - 0.5, 0.5 * Ny - 0.5, 0.5 * Nz - 0.5), np.float32(stl_size), rotation=ROT), flag=TYPE_SX)
        body = (ref.get_global("flags") & 0x40) != 0
        rho = np.where(body, np.float32(1.0), rho); u = [np.where(body, np.float32(0.0), a) for a in u]; flags = np.where(body, np.uint8(TYPE_SX), flags)
    load_scenario(ref, rho, u, flags)
    if mb:
        z, uy = int(mb.split(",")[0]), np.float32(float(mb.split(",")[1]))
        ref.run(steps // 2)
        rf = ref.fields()  # lbm.u.read_from_device() / lbm.flags.read_from_device(): the host copies now hold the device's view
        new_uy = rf[2].copy()
        plane = (rf[4][z] & 3) == 1
        new_uy[z][plane] = uy
        for a, arr in enumerate((rf[1], new_uy, rf[3])): ref.set_global("u", arr, a)
        ref._communicate("ruf")  # (write_to_device() re-sends whole domains, halos included, from the host copies read above)
        ref.update_moving_boundaries()
        ref.run(steps - steps // 2)
    else:
        ref.run(steps)
    want = ref.fields()
    if mb: assert np.any((want[4] & 3) == 3), "no TYPE_MS cells in the scene"
    for name, a, b in zip(("rho", "ux", "uy", "uz", "flags"), got, want):
        a = a.view(np.uint32) if a.dtype == np.float32 else a
        b = b.view(np.uint32) if b.dtype == np.float32 else b
        assert np.array_equal(a, b), f"{name} differs from the oracle in {int(np.sum(a != b))} cells"
    if stl:
        assert int(np.sum(want[4] == TYPE_SX)) > 100, "the voxelised body is missing"
        ref.update_force_field()
        F = np.stack([ref.get_global("F", a) for a in range(3)])
        assert np.array_equal(got_ff[:3 * N].view(np.uint32), F.ravel().view(np.uint32)), "F differs from the oracle"
        c = (np.float32(0.5) * np.float32(Nx) - np.float32(0.5), np.float32(0.5) * np.float32(Ny) - np.float32(0.5), np.float32(0.5) * np.float32(Nz) - np.float32(0.5))
        sums = np.concatenate([ref.object_sum(1, TYPE_SX), ref.object_sum(0, TYPE_SX), ref.object_sum(2, TYPE_SX, center=c)]).astype(np.float32)
        assert np.array_equal(got_ff[3 * N:].view(np.uint32), sums.view(np.uint32)), (got_ff[3 * N:], sums)
        assert np.any(sums[:3] != 0)


@pytest.fixture(scope="module")
def emul_so():
    subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "tests", "emul")], check=True, capture_output=True)
    return EMUL_SO


@pytest.mark.parametrize("case", CASES, ids=case_id)
def test_cpp_host_scene_matches_oracle_emulated(emul_so, case, tmp_path):
    run_case(case, emul_so, "emul", tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("case", GPU_CASES, ids=case_id)
def test_cpp_host_scene_matches_oracle_gpu(case, tmp_path):
    assert os.path.exists(PRODUCT_SO), "libfx3d_cuda.so is not built"
    run_case(case, PRODUCT_SO, "cuda", tmp_path)
