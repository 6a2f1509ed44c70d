"""Product host logic (fluidx3d_b200/lbm.py) + product kernel sources, compiled against tests/emul/cuda_emul.hpp and run
as OS threads on the CPU, against the oracle. This is a development aid for a GPU-less container: it checks addressing,
warp-shuffle hand-over, pack/unpack and operation order of the kernels; the GPU parity tests (-m gpu) are the gate."""
import os
import subprocess
import numpy as np
import pytest
from helpers import (ROOT, OracleBackend, HostSim, scenario, load_scenario, FP32, FP16S, FP16C, SRT, TRT)
from fluidx3d_b200 import capi
from fluidx3d_b200 import lbm as lbm_mod
from fluidx3d_b200.lbm import LBM

lbm_mod.VERBOSE = False
EMUL_SO = os.path.join(ROOT, "tests", "_build", "libfx3d_emul.so")


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "tests", "emul")], check=True, capture_output=True)
    return capi.Lib(EMUL_SO)


def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def product(lib, v, dims, D, steps, f, variant, nu=0.05, seed=3):
    Q, coll, st, feat = v
    lib.set_kernel_variant(variant)
    sim = LBM(*dims, nu, *f, Dx=D[0], Dy=D[1], Dz=D[2], velocity_set=Q, collision=coll, storage=st, features=feat, lib=lib)
    rho, u, flags = scenario(sim.Nx, sim.Ny, sim.Nz, seed=seed, eq_frac=0.03 if feat & 2 else 0.0)
    sim.rho.set_global(rho); [sim.u.set_global(u[a], a) for a in range(3)]; sim.flags.set_global(flags)
    sim.run(steps)
    for m in (sim.rho, sim.u, sim.flags): m.read_from_device()
    out = (sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2), sim.flags.get_global())
    sim.close()
    lib.set_kernel_variant(0)
    return out


def oracle(v, dims, D, steps, f, nu=0.05, seed=3):
    Q, coll, st, feat = v
    sim = HostSim(OracleBackend(Q, coll, st, feat), *dims, *D, nu=nu, fx=f[0], fy=f[1], fz=f[2])
    load_scenario(sim, *scenario(sim.Nx, sim.Ny, sim.Nz, seed=seed, eq_frac=0.03 if feat & 2 else 0.0))
    sim.run(steps)
    return sim.fields()


CASES = [((19, SRT, FP32, 0), (16, 6, 4), (1, 1, 1)), ((19, SRT, FP16S, 0), (16, 6, 4), (1, 1, 1)), ((19, SRT, FP16C, 0), (8, 6, 5), (1, 1, 1)),
         ((19, TRT, FP32, 3), (16, 6, 6), (2, 1, 2)), ((27, TRT, FP32, 3), (8, 6, 5), (1, 1, 1)), ((27, SRT, FP16S, 0), (16, 6, 6), (2, 1, 2)),
         ((19, SRT, FP16S, 0), (12, 6, 6), (2, 2, 1)), ((19, SRT, FP32, 0), (7, 6, 5), (1, 1, 1))]


@pytest.mark.parametrize("v,dims,D", CASES, ids=[f"q{c[0][0]}c{c[0][1]}s{c[0][2]}f{c[0][3]}-{'x'.join(map(str, c[1]))}-d{''.join(map(str, c[2]))}" for c in CASES])
@pytest.mark.parametrize("variant", [1, 2, 4, 8, 32], ids=["general", "vector2", "vector4", "pipelined", "occupancy"])
def test_emulated_kernels_match_oracle(emul, v, dims, D, variant):
    f = (1e-4, -2e-4, 3e-4) if v[3] & 1 else (0.0, 0.0, 0.0)
    for steps in (1, 4):
        got, want = product(emul, v, dims, D, steps, f, variant), oracle(v, dims, D, steps, f)
        for a, b in zip(got, want):
            assert np.array_equal(bits(a), bits(b))


TMA_CASES = [((19, SRT, FP16S, 0), (32, 16, 4), (1, 1, 1)), ((19, SRT, FP32, 0), (64, 8, 3), (1, 1, 1)), ((19, TRT, FP16C, 3), (16, 32, 3), (1, 1, 1)),
             ((27, TRT, FP32, 3), (32, 16, 3), (1, 1, 1)), ((27, SRT, FP16S, 1), (32, 32, 4), (1, 2, 2)), ((19, SRT, FP16S, 2), (64, 16, 6), (1, 2, 1)),
             ((19, SRT, FP16S, 0), (512, 3, 3), (1, 1, 1)), ((19, TRT, FP32, 3), (512, 2, 4), (1, 1, 2)), ((27, SRT, FP16C, 0), (512, 2, 2), (1, 1, 1)),  # one-row tiles: compile-time copy lists
             # long runs of tiles per block (more than stages + groups): every stage is refilled several times, by a group other than the one that reads it
             ((19, SRT, FP16S, 0), (16, 32, 40), (1, 1, 1)), ((19, SRT, FP32, 2), (32, 16, 48), (1, 1, 1)), ((19, SRT, FP16S, 0), (512, 2, 20), (1, 1, 1))]


@pytest.mark.parametrize("v,dims,D", TMA_CASES, ids=[f"q{c[0][0]}c{c[0][1]}s{c[0][2]}f{c[0][3]}-{'x'.join(map(str, c[1]))}-d{''.join(map(str, c[2]))}" for c in TMA_CASES])
def test_emulated_bulk_copy_kernel_matches_oracle(emul, v, dims, D):
    """the bulk-copy (TMA) form of stream_collide on grids where it is eligible (tile = whole rows): row buffers in shared
    memory, shifted reads/writes of the periodic row, stage recycling, two block barriers per tile"""
    f = (1e-4, -2e-4, 3e-4) if v[3] & 1 else (0.0, 0.0, 0.0)
    before = emul.kernel_kind_counts()
    for steps in (1, 2, 5):
        got, want = product(emul, v, dims, D, steps, f, 16), oracle(v, dims, D, steps, f)
        for a, b in zip(got, want):
            assert np.array_equal(bits(a), bits(b))
    after = emul.kernel_kind_counts()
    assert after[3] > before[3] and after[:3] == before[:3], "the whole-row bulk-copy kernel must be the one that ran"


SEG_CASES = [((19, SRT, FP16S, 0), (1024, 2, 2), (1, 1, 1)), ((19, SRT, FP16S, 0), (256, 4, 3), (2, 1, 1)), ((19, TRT, FP16S, 3), (1024, 2, 2), (2, 1, 1)), ((19, TRT, FP32, 3), (256, 8, 4), (2, 2, 1)),
             ((27, SRT, FP16C, 2), (1024, 2, 2), (1, 1, 1)), ((19, SRT, FP32, 1), (1024, 3, 2), (1, 1, 2)), ((27, TRT, FP16S, 3), (512, 8, 2), (4, 2, 1))]


@pytest.mark.parametrize("v,dims,D", SEG_CASES, ids=[f"q{c[0][0]}c{c[0][1]}s{c[0][2]}f{c[0][3]}-{'x'.join(map(str, c[1]))}-d{''.join(map(str, c[2]))}" for c in SEG_CASES])
def test_emulated_hybrid_kernel_matches_oracle(emul, v, dims, D):
    """the default choice for row segments and x-decomposed domains: bulk loads into padded row buffers, stream-out straight
    from registers with the vector kernel's ownership rules (shuffles inside a warp, scalar stores at warp and segment ends)"""
    f = (1e-4, -2e-4, 3e-4) if v[3] & 1 else (0.0, 0.0, 0.0)
    before = emul.kernel_kind_counts()
    for steps in (1, 2, 5):
        got, want = product(emul, v, dims, D, steps, f, 0), oracle(v, dims, D, steps, f)
        for a, b in zip(got, want):
            assert np.array_equal(bits(a), bits(b))
    after = emul.kernel_kind_counts()
    if (dims[0] // D[0]) != 512:
        assert after[5] > before[5] and after[:5] == before[:5], "the hybrid kernel must be the one that ran"
    else:  # x-decomposed domains whose rows are exactly one tile: the whole-row kernel takes them (halo cells in the row-buffer pads)
        assert after[3] > before[3] and after[:3] == before[:3] and after[5] == before[5], "the whole-row bulk-copy kernel must be the one that ran"


XROW_CASES = [((27, TRT, FP32, 3), (1024, 4, 6), (2, 1, 2)),  # D3Q27 FP32: 16-byte pads, and the one direction pair whose neighbour side reaches the LEFT halo cell
              ((19, SRT, FP16S, 1), (1024, 4, 4), (2, 1, 2)), ((19, TRT, FP32, 3), (1024, 2, 3), (2, 2, 1)), ((27, SRT, FP16C, 2), (1024, 2, 2), (2, 1, 1)),  # x halos: one-row tiles
              ((19, SRT, FP32, 0), (32, 32, 8), (1, 2, 2)), ((19, TRT, FP16C, 11), (32, 32, 6), (1, 2, 1)), ((27, TRT, FP32, 3), (64, 16, 6), (1, 2, 2)),
              ((19, SRT, FP16S, 24), (64, 32, 8), (1, 2, 2)), ((19, SRT, FP16S, 0), (512, 2, 6), (1, 2, 2))]  # (rows per domain: a multiple of the rows per tile)


@pytest.mark.parametrize("v,dims,D", XROW_CASES, ids=[f"q{c[0][0]}c{c[0][1]}s{c[0][2]}f{c[0][3]}-{'x'.join(map(str, c[1]))}-d{''.join(map(str, c[2]))}" for c in XROW_CASES])
def test_emulated_row_kernel_with_halos_and_fused_delivery(emul, v, dims, D):
    """the whole-row bulk-copy kernel on decomposed domains: x halos live in the pads of the row buffers (no periodic wrap), and the y/z
    part of the halo exchange is fused into the kernel -- every stored row goes to the domain that reads it next (LBM.do_time_step picks
    fx3d_stream_collide_fused where fx3d_fused_halo_supported says so)"""
    f = (1e-4, -2e-4, 3e-4) if v[3] & 1 else (0.0, 0.0, 0.0)
    before = emul.kernel_kind_counts()
    for steps in (1, 2, 5):
        got, want = product(emul, v, dims, D, steps, f, 0), oracle(v, dims, D, steps, f)
        for a, b in zip(got, want):
            assert np.array_equal(bits(a), bits(b))
    after = emul.kernel_kind_counts()
    assert after[3] > before[3] and all(after[k] == before[k] for k in (0, 1, 2, 4, 5, 6)), "the whole-row bulk-copy kernel must be the one that ran"


SG_CASES = [((19, SRT, FP32, 8), (9, 7, 5), (1, 1, 1)), ((19, TRT, FP16S, 11), (12, 6, 6), (2, 1, 2)), ((27, SRT, FP16C, 8), (10, 6, 4), (1, 1, 1)),  # general kernel
            ((19, SRT, FP16S, 8), (32, 16, 4), (1, 1, 1)), ((19, TRT, FP32, 11), (64, 8, 4), (1, 2, 1)), ((27, TRT, FP16C, 9), (32, 16, 3), (1, 1, 1)),    # whole-row bulk-copy kernel
            ((19, SRT, FP16S, 8), (1024, 2, 2), (1, 1, 1)), ((19, TRT, FP32, 11), (256, 4, 3), (2, 1, 1)), ((27, SRT, FP16C, 8), (1024, 2, 2), (2, 1, 1))]   # row segments / x halos: hybrid kernel


@pytest.mark.parametrize("v,dims,D", SG_CASES, ids=[f"q{c[0][0]}c{c[0][1]}s{c[0][2]}f{c[0][3]}-{'x'.join(map(str, c[1]))}-d{''.join(map(str, c[2]))}" for c in SG_CASES])
def test_emulated_subgrid_matches_oracle(emul, v, dims, D):
    """SUBGRID (feature bit 3, Smagorinsky-Lilly): per-cell relaxation rate from the non-equilibrium stress tensor, in scalar
    arithmetic (general kernel) and in packed lanes at the FP16S working scale (bulk-copy kernel); low viscosity so it matters"""
    f = (1e-4, -2e-4, 3e-4) if v[3] & 1 else (0.0, 0.0, 0.0)
    for steps in (1, 4):
        got, want = product(emul, v, dims, D, steps, f, 0, nu=0.002), oracle(v, dims, D, steps, f, nu=0.002)
        for a, b in zip(got, want):
            assert np.array_equal(bits(a), bits(b))


MB_CASES = [((19, SRT, FP32, 16), (9, 7, 5), (1, 1, 1)), ((19, TRT, FP16S, 19), (12, 6, 6), (2, 1, 2)), ((27, SRT, FP16C, 18), (10, 6, 4), (1, 1, 1)),
            ((19, SRT, FP16S, 24), (16, 8, 4), (1, 2, 1)),
            # whole-row tiles: the bulk-copy kernel with the rare-lane path in collide_tile
            ((19, SRT, FP16S, 16), (32, 16, 4), (1, 1, 1)), ((19, TRT, FP32, 19), (64, 8, 4), (1, 2, 1)), ((27, SRT, FP16C, 18), (32, 16, 3), (1, 1, 1)),
            ((19, SRT, FP16S, 24), (64, 8, 4), (1, 1, 1)),
            # row segments / x halos: the hybrid kernel
            ((19, SRT, FP16S, 16), (1024, 2, 2), (1, 1, 1)), ((19, TRT, FP32, 19), (256, 4, 3), (2, 1, 1)), ((27, SRT, FP16C, 26), (1024, 2, 2), (2, 1, 1))]


@pytest.mark.parametrize("v,dims,D", MB_CASES, ids=[f"q{c[0][0]}c{c[0][1]}s{c[0][2]}f{c[0][3]}-{'x'.join(map(str, c[1]))}-d{''.join(map(str, c[2]))}" for c in MB_CASES])
def test_emulated_moving_boundaries_match_oracle(emul, v, dims, D):
    """MOVING_BOUNDARIES (feature bit 4): TYPE_MS marking in initialize, the Dirichlet correction in stream_collide and
    update_fields, and update_moving_boundaries() after boundary velocities changed -- general kernel, against the oracle"""
    from fluidx3d_b200.lbm import LBM
    Q, coll, st, feat = v
    f = (1e-4, -2e-4, 3e-4) if feat & 1 else (0.0, 0.0, 0.0)
    sim = LBM(*dims, 0.05, *f, Dx=D[0], Dy=D[1], Dz=D[2], velocity_set=Q, collision=coll, storage=st, features=feat, lib=emul)
    ref = HostSim(OracleBackend(Q, coll, st, feat), *dims, *D, nu=0.05, fx=f[0], fy=f[1], fz=f[2])
    rho, u, flags = scenario(sim.Nx, sim.Ny, sim.Nz, seed=7, eq_frac=0.03 if feat & 2 else 0.0)
    sim.rho.set_global(rho); [sim.u.set_global(u[a], a) for a in range(3)]; sim.flags.set_global(flags)
    load_scenario(ref, rho, u, flags)
    sim.run(3); ref.run(3)
    # two boundaries change speed: one stops, one starts moving
    solid = np.argwhere((flags & 3) == 1)
    u2 = [a.copy() for a in u]
    (z0, y0, x0), (z1, y1, x1) = solid[0], solid[1]
    for a in range(3): u2[a][z0, y0, x0] = 0.0
    u2[0][z1, y1, x1] = np.float32(0.02)
    for m in (sim.rho, sim.u, sim.flags): m.read_from_device()
    cur_flags = sim.flags.get_global()
    [sim.u.set_global(np.where((cur_flags & 3) == 1, u2[a], sim.u.get_global(a)), a) for a in range(3)]
    sim.u.write_to_device()
    if sim.get_D() > 1: sim.communicate_rho_u_flags()  # the neighbours' halo copies of u, as in ref._communicate below
    sim.update_moving_boundaries()
    rf = ref.fields()
    for a in range(3): ref.set_global("u", np.where((rf[4] & 3) == 1, u2[a], rf[1 + a]), a)
    ref._communicate("ruf")
    ref.update_moving_boundaries()
    sim.run(3); ref.run(3)
    for m in (sim.rho, sim.u, sim.flags): m.read_from_device()
    got = (sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2), sim.flags.get_global())
    want = ref.fields()
    assert np.any((want[4] & 3) == 3), "the scene has no TYPE_MS cells"
    for a, b in zip(got, want):
        assert np.array_equal(bits(a), bits(b))
    sim.close()


@pytest.mark.parametrize("variant", [1, 4, 8], ids=["general", "vector4", "pipelined"])
def test_shell_plus_interior_equals_all(emul, variant):
    # FX3D_REGION_SHELL followed by FX3D_REGION_INTERIOR must cover every non-halo cell exactly once
    v, dims, D, f = (19, SRT, FP16S, 0), (24, 10, 8), (2, 2, 2), (0.0, 0.0, 0.0)
    emul.set_kernel_variant(variant)
    outs = []
    for overlap in (False, True):
        sim = LBM(*dims, 0.05, Dx=D[0], Dy=D[1], Dz=D[2], velocity_set=v[0], collision=v[1], storage=v[2], features=v[3], lib=emul, overlap=overlap)
        rho, u, flags = scenario(*dims, seed=5)
        sim.rho.set_global(rho); [sim.u.set_global(u[a], a) for a in range(3)]; sim.flags.set_global(flags)
        sim.run(3)
        for m in (sim.rho, sim.u): m.read_from_device()
        outs.append((sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2)))
        sim.close()
    emul.set_kernel_variant(0)
    for a, b in zip(*outs):
        assert np.array_equal(bits(a), bits(b))
    want = oracle(v, dims, D, 3, f, seed=5)
    for a, b in zip(outs[1], want):
        assert np.array_equal(bits(a), bits(b))


OVERLAP_CASES = [((27, SRT, FP32, 0), (24, 6, 6), (2, 1, 1)),      # SHELL on the vector kernel (K=4), INTERIOR on the ring kernel (K=2): the x split must not follow K
                 ((19, SRT, FP32, 8), (1040, 2, 2), (2, 1, 1)),    # SUBGRID: general kernel for one region, hybrid kernel for the other
                 ((19, SRT, FP16S, 8), (128, 16, 8), (1, 2, 2)),   # SUBGRID: z slabs eligible for the bulk-copy kernel, one-row y slabs not (per-region fallback)
                 ((19, TRT, FP16S, 19), (24, 10, 8), (2, 2, 2)),   # MOVING_BOUNDARIES + force + TYPE_E
                 ((19, SRT, FP16C, 0), (22, 6, 6), (2, 1, 2)),     # row length divisible by 2 only
                 ((27, TRT, FP16S, 3), (18, 5, 7), (2, 1, 1))]     # odd row length: one-cell x shell


@pytest.mark.parametrize("v,dims,D", OVERLAP_CASES, ids=[f"q{c[0][0]}c{c[0][1]}s{c[0][2]}f{c[0][3]}-{'x'.join(map(str, c[1]))}-d{''.join(map(str, c[2]))}" for c in OVERLAP_CASES])
def test_shell_plus_interior_default_dispatch(emul, v, dims, D):
    """overlap mode with the library's own kernel choice (variant 0): the SHELL and INTERIOR passes may pick different kernel
    forms with different cells per thread, and still every non-halo cell must be advanced exactly once (round-1 advisor finding)"""
    f = (1e-4, -2e-4, 3e-4) if v[3] & 1 else (0.0, 0.0, 0.0)
    emul.set_kernel_variant(0)
    sim = LBM(*dims, 0.05, *f, Dx=D[0], Dy=D[1], Dz=D[2], velocity_set=v[0], collision=v[1], storage=v[2], features=v[3], lib=emul, overlap=True)
    rho, u, flags = scenario(sim.Nx, sim.Ny, sim.Nz, seed=3, eq_frac=0.03 if v[3] & 2 else 0.0)
    sim.rho.set_global(rho); [sim.u.set_global(u[a], a) for a in range(3)]; sim.flags.set_global(flags)
    sim.run(3)
    for m in (sim.rho, sim.u, sim.flags): m.read_from_device()
    got = (sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2), sim.flags.get_global())
    sim.close()
    want = oracle(v, dims, D, 3, f)
    for a, b in zip(got, want):
        assert np.array_equal(bits(a), bits(b))


def test_library_exports_every_declared_symbol():
    # the product library itself (built by nvcc) must load without a GPU and export everything include/fx3d.h declares
    import re
    hdr = open(os.path.join(ROOT, "include", "fx3d.h")).read()
    declared = set(re.findall(r"\b(fx3d_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    if not os.path.exists(capi.DEFAULT_LIB):
        subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "fluidx3d_b200", "csrc")], check=True, capture_output=True)
    capi.Lib()  # raises if a symbol is missing


def test_relaxation_rate_matches_oracle_round_trip():
    lib = capi.Lib()
    orc = OracleBackend()
    for nu in (1.0, 1.0 / 6.0, 0.01, 0.02, 0.05, 1e-3, 3.3e-5, 0.0062, 7.5e-4, 12.5):
        assert np.float32(lib.relaxation_rate(nu)) == np.float32(orc.w_from_nu(nu))


def test_constructor_checks():
    lib = capi.Lib(EMUL_SO) if os.path.exists(EMUL_SO) else None
    if lib is None:
        pytest.skip("emulation library not built")
    with pytest.raises(ValueError): LBM(0, 4, 4, 0.1, lib=lib)
    with pytest.raises(ValueError): LBM(4, 4, 4, 0.0, lib=lib)
    with pytest.raises(ValueError): LBM(4, 4, 4, -1.0, lib=lib)
    with pytest.raises(ValueError): LBM(4, 4, 4, 0.1, 1e-3, 0.0, 0.0, lib=lib)  # force without VOLUME_FORCE
    with pytest.raises(ValueError): LBM(4, 4, 4, 0.1, Dx=0, lib=lib)


def test_c_abi_error_behaviour(emul):
    """every entry point returns a status code and leaves a message; nothing exits (the C++ wrapper maps them to print_error)"""
    import ctypes as C
    from fluidx3d_b200.capi import Fx3dError
    for bad in (3, 5, 7, -1, 64):
        with pytest.raises(Fx3dError, match="variant must be"): emul.set_kernel_variant(bad)
    for ok in (0, 1, 2, 4, 8, 16, 32, 0): emul.set_kernel_variant(ok)
    with pytest.raises(Fx3dError, match="reserve"): emul.set_interior_reserve(-1)
    emul.set_interior_reserve(8)
    n = C.c_uint64(0)
    with pytest.raises(Fx3dError, match="kind"): emul.stream_collide_launches(9, C.byref(n))
    sim = LBM(8, 4, 4, 0.1, lib=emul)  # no MOVING_BOUNDARIES
    with pytest.raises(ValueError): sim.update_moving_boundaries()
    (_, dom), = sim.local_domains()
    with pytest.raises(Fx3dError, match="MOVING_BOUNDARIES"): emul.update_moving_boundaries(C.byref(dom.lat), dom.stream)
    sim.close()
