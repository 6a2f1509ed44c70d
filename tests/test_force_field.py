"""FORCE_FIELD (src/kernel.cpp:1497-1503,1821-1827,1873-1959,2173-2196; src/lbm.cpp:206-239,986-1016):
the per-cell force F in stream_collide / update_fields (with VOLUME_FORCE), update_force_field (boundary forces on TYPE_S cells),
the object_center_of_mass / object_force / object_torque sums, and the F halo exchange.

  * oracle == oracle/_ref (the reference's own kernels compiled natively) bit for bit: fields, raw DDFs and F
  * golden vectors generated from oracle/_ref (tests/golden/ff_*.npz, oracle/make_golden.py) -- also where _ref is absent
  * product (Python host + kernels) == oracle bit for bit: in emulation on the CPU, on the B200 with -m gpu
  * the object sums: the reference adds work-group partial sums with floating-point atomics (any order); the oracle and the product
    fix the order (ascending), so they agree bit for bit, and both agree with a float64 sum within rounding
  * a physical check: the momentum-exchange force on a sphere in uniform flow points downstream"""
import glob
import os
import subprocess
import numpy as np
import pytest
from helpers import (ROOT, OracleBackend, RefBackend, HostSim, scenario, load_scenario, ref_available, FP32, FP16S, FP16C, SRT, TRT, TYPE_S, TYPE_E)
from fluidx3d_b200 import capi
from fluidx3d_b200 import lbm as lbm_mod
from fluidx3d_b200.lbm import LBM

lbm_mod.VERBOSE = False
EMUL_SO = os.path.join(ROOT, "tests", "_build", "libfx3d_emul.so")
TYPE_X = 0x40
VARIANTS = [(19, SRT, FP32, 33), (19, TRT, FP16S, 35), (27, SRT, FP16C, 33), (19, SRT, FP32, 34)]  # bit5 FORCE_FIELD with / without VOLUME_FORCE


def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def force_of(feat):
    return (1e-4, -2e-4, 3e-4) if feat & 1 else (0.0, 0.0, 0.0)


def scene(Nx, Ny, Nz, feat, seed):
    rho, u, flags = scenario(Nx, Ny, Nz, seed=seed, eq_frac=0.03 if feat & 2 else 0.0)
    idx = np.arange(flags.size).reshape(flags.shape)
    flags[((flags & 3) == TYPE_S) & (idx % 2 == 0)] |= np.uint8(TYPE_X)  # half of the solids carry a marker
    rng = np.random.default_rng(seed)
    F = [(2e-3 * (rng.random((Nz, Ny, Nx), dtype=np.float32) - 0.5)).astype(np.float32) for _ in range(3)]
    return rho, u, flags, F


def run_host(cls, v, dims, D, steps, seed):
    Q, coll, st, feat = v
    f = force_of(feat)
    sim = HostSim(cls(Q, coll, st, feat), *dims, *D, nu=0.04, fx=f[0], fy=f[1], fz=f[2])
    rho, u, flags, F = scene(sim.Nx, sim.Ny, sim.Nz, feat, seed)
    load_scenario(sim, rho, u, flags)
    for a in range(3): sim.set_global("F", F[a], a)
    sim.run(steps)
    out = list(sim.fields())
    sim.update_force_field()
    out += [sim.get_global("F", a) for a in range(3)] + [d.fi.copy() for d in sim.dom]
    sim.run(2)  # the boundary forces now act on the solids' own cells only (F of fluid cells is untouched): the run continues with them
    out += list(sim.fields())
    return sim, out


@pytest.mark.parametrize("v", VARIANTS, ids=[f"q{v[0]}c{v[1]}s{v[2]}f{v[3]}" for v in VARIANTS])
def test_oracle_equals_reference_device_code_force_field(v):
    if not ref_available(*v):
        pytest.skip("oracle/_ref variant not built")
    for dims, D, steps, seed in [((9, 7, 5), (1, 1, 1), 3, 1), ((12, 8, 6), (2, 1, 1), 4, 2), ((8, 8, 8), (2, 2, 2), 5, 3), ((6, 10, 8), (1, 2, 2), 2, 4)]:
        _, a = run_host(OracleBackend, v, dims, D, steps, seed)
        _, r = run_host(RefBackend, v, dims, D, steps, seed)
        for x, y in zip(a, r):
            assert np.array_equal(bits(x), bits(y)), (v, dims, D, steps)


GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ff_*.npz")))


def test_force_field_golden_files_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(f)[:-4] for f in GOLDEN])
def test_oracle_reproduces_force_field_golden(path):
    g = np.load(path)
    Q, coll, st, feat, Nx, Ny, Nz, Dx, Dy, Dz, steps, seed = (int(x) for x in g["meta"])
    _, out = run_host(OracleBackend, (Q, coll, st, feat), (Nx, Ny, Nz), (Dx, Dy, Dz), steps, seed)
    for k, x in enumerate(out):
        assert np.array_equal(bits(np.asarray(x)), bits(g[f"out{k}"])), (path, k)


def test_object_sums_fixed_order_and_tolerance():
    """oracle object sums: deterministic, equal to a float64 sum within float rounding, cell count exact; group size changes only the rounding"""
    v, dims = (19, SRT, FP32, 34), (20, 12, 10)
    sim, _ = run_host(OracleBackend, v, dims, (1, 1, 1), 4, 5)
    d = sim.dom[0]
    N = d.flags.size
    marker = TYPE_S | TYPE_X
    sel = d.flags == marker
    assert sel.sum() > 10
    force = sim.b.object_sum(1, d.F, d.flags, marker)
    want = np.array([d.F[a * N:(a + 1) * N][sel].astype(np.float64).sum() for a in range(3)])
    assert np.allclose(force[:3], want, rtol=1e-5, atol=1e-9)
    assert np.array_equal(force, sim.b.object_sum(1, d.F, d.flags, marker))
    assert np.allclose(sim.b.object_sum(1, d.F, d.flags, marker, group=32)[:3], want, rtol=1e-5, atol=1e-9)
    com = sim.b.object_sum(0, None, d.flags, marker)
    assert int(com[3:4].view(np.uint32)[0]) == int(sel.sum())
    zz, yy, xx = np.nonzero(sel.reshape(dims[2], dims[1], dims[0]))
    pos = np.array([(xx + 0.5 - 0.5 * dims[0]).sum(), (yy + 0.5 - 0.5 * dims[1]).sum(), (zz + 0.5 - 0.5 * dims[2]).sum()])
    assert np.allclose(com[:3], pos, rtol=1e-5, atol=1e-3)
    c = (0.5, -1.0, 2.0)
    tq = sim.b.object_sum(2, d.F, d.flags, marker, center=c)
    r = np.stack([xx + 0.5 - 0.5 * dims[0] - c[0], yy + 0.5 - 0.5 * dims[1] - c[1], zz + 0.5 - 0.5 * dims[2] - c[2]], axis=1)
    Fv = np.stack([d.F[a * N:(a + 1) * N][sel].astype(np.float64) for a in range(3)], axis=1)
    assert np.allclose(tq[:3], np.cross(r, Fv).sum(axis=0), rtol=1e-4, atol=1e-7)


def test_sphere_drag_points_downstream():
    """physics: a sphere in a box driven through TYPE_E faces with u=(0.05,0,0); the momentum-exchange force on it points along +x"""
    n, R, u0 = 32, 5.0, 0.05
    zz, yy, xx = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    flags = np.zeros((n, n, n), np.uint8)
    for sl in [np.s_[0], np.s_[-1], np.s_[:, 0], np.s_[:, -1], np.s_[:, :, 0], np.s_[:, :, -1]]: flags[sl] = TYPE_E
    sphere = (xx - n / 2) ** 2 + (yy - n / 2) ** 2 + (zz - n / 2) ** 2 <= R * R
    flags[sphere] = TYPE_S | TYPE_X
    sim = HostSim(OracleBackend(19, SRT, FP32, 34), n, n, n, nu=0.1)
    sim.set_global("flags", flags)
    sim.set_global("u", np.where(sphere, 0.0, u0).astype(np.float32), 0)
    sim.run(300)
    force = sim.object_sum(1, TYPE_S | TYPE_X)
    assert force[0] > 1e-4 and abs(force[1]) < 0.05 * force[0] and abs(force[2]) < 0.05 * force[0], force
    com = sim.object_sum(0, TYPE_S | TYPE_X)
    assert np.allclose(com, [0.5, 0.5, 0.5], atol=1e-4), com  # the sphere is centred on cell n/2: position n/2 + 0.5 - n/2


# ---------------- product: Python host + kernels ----------------
def run_product(lib, v, dims, D, steps, seed, devices=None):
    Q, coll, st, feat = v
    f = force_of(feat)
    sim = LBM(*dims, 0.04, *f, Dx=D[0], Dy=D[1], Dz=D[2], velocity_set=Q, collision=coll, storage=st, features=feat, lib=lib, devices=devices)
    rho, u, flags, F = scene(sim.Nx, sim.Ny, sim.Nz, feat, seed)
    sim.rho.set_global(rho); [sim.u.set_global(u[a], a) for a in range(3)]; sim.flags.set_global(flags); [sim.F.set_global(F[a], a) for a in range(3)]
    sim.run(steps)
    for m in (sim.rho, sim.u, sim.flags): m.read_from_device()
    out = [sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2), sim.flags.get_global()]
    sim.update_force_field()
    sim.F.read_from_device()
    out += [sim.F.get_global(a) for a in range(3)]
    marker = TYPE_S | TYPE_X
    sums = [sim.object_force(marker), sim.object_center_of_mass(marker), sim.object_torque((0.5, -1.0, 2.0), marker)]
    sim.run(2)
    for m in (sim.rho, sim.u, sim.flags): m.read_from_device()
    out += [sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2), sim.flags.get_global()]
    sim.close()
    return out, sums


def check_product(lib, v, dims, D, steps, seed, devices=None):
    got, sums = run_product(lib, v, dims, D, steps, seed, devices)
    ref, want = run_host(OracleBackend, v, dims, D, steps, seed)
    want = want[:8] + want[8 + len(ref.dom):]  # without the raw DDF buffers (the product's layout is private)
    for k, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(bits(a), bits(b)), (v, dims, D, k, int(np.sum(bits(a) != bits(b))))
    # object sums against the oracle's fixed-order sums. update_force_field ran at the same step in both (t = steps), F is identical (checked above)
    marker = TYPE_S | TYPE_X
    ref.t_last_force_field = ref.t  # F as of the update above; the two extra steps did not touch it
    for kind, s in zip((1, 0, 2), sums):
        w = ref.object_sum(kind, marker, center=(0.5, -1.0, 2.0))
        assert np.array_equal(bits(np.asarray(s, np.float32)), bits(np.asarray(w, np.float32))), (kind, s, w)


PRODUCT_CASES = [((19, SRT, FP32, 33), (12, 8, 6), (1, 1, 1), 4, 2), ((19, TRT, FP16S, 35), (12, 8, 6), (2, 1, 2), 4, 3), ((27, SRT, FP16C, 33), (8, 8, 8), (1, 2, 1), 3, 4),
                 ((19, SRT, FP32, 34), (16, 8, 6), (1, 1, 1), 4, 5),   # no VOLUME_FORCE: stream_collide keeps its fast kernels, F only receives the boundary forces
                 ((19, SRT, FP16S, 34), (32, 16, 4), (1, 2, 2), 4, 6)]  # whole-row tiles with fused halo delivery
IDS = [f"q{c[0][0]}c{c[0][1]}s{c[0][2]}f{c[0][3]}-{'x'.join(map(str, c[1]))}-d{''.join(map(str, c[2]))}" for c in PRODUCT_CASES]


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "tests", "emul")], check=True, capture_output=True)
    return capi.Lib(EMUL_SO)


@pytest.mark.parametrize("v,dims,D,steps,seed", PRODUCT_CASES, ids=IDS)
def test_emulated_force_field_matches_oracle(emul, v, dims, D, steps, seed):
    check_product(emul, v, dims, D, steps, seed)


GPU_CASES = PRODUCT_CASES + [((19, SRT, FP32, 34), (128, 64, 48), (1, 1, 1), 10, 7), ((19, TRT, FP32, 35), (64, 48, 32), (2, 2, 1), 6, 8), ((27, SRT, FP16S, 34), (64, 64, 32), (1, 1, 2), 6, 9)]


@pytest.mark.gpu
@pytest.mark.parametrize("v,dims,D,steps,seed", GPU_CASES, ids=[f"q{c[0][0]}c{c[0][1]}s{c[0][2]}f{c[0][3]}-{'x'.join(map(str, c[1]))}-d{''.join(map(str, c[2]))}" for c in GPU_CASES])
def test_force_field_bit_exact_gpu(v, dims, D, steps, seed):
    check_product(capi.lib(), v, dims, D, steps, seed, devices=[0] * (D[0] * D[1] * D[2]))
