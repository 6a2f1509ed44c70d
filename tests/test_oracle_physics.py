"""Size-independent properties of the path, checked on the CPU oracle: mass conservation, decomposition invariance
(D>1 bit-identical to D=1, cf. reference changelog README.md:154), implicit bounce-back turn-around, Poiseuille
(src/setup.cpp:84-144, expected L2 error 2-5 %) and Taylor-Green decay."""
import numpy as np
import pytest
from helpers import (OracleBackend, HostSim, scenario, load_scenario, FP32, FP16S, FP16C, SRT, TRT, VOLUME_FORCE, TYPE_S)


def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


@pytest.mark.parametrize("Q,coll,st", [(19, SRT, FP32), (27, TRT, FP32), (19, SRT, FP16S), (19, TRT, FP16C)])
def test_decomposition_is_bit_identical(Q, coll, st):
    outs = []
    for D in [(1, 1, 1), (2, 1, 1), (1, 2, 1), (1, 1, 2), (2, 2, 1), (2, 2, 2), (4, 1, 2)]:
        sim = HostSim(OracleBackend(Q, coll, st, 0), 16, 12, 8, *D, nu=0.04)
        load_scenario(sim, *scenario(16, 12, 8, seed=9))
        sim.run(7)
        outs.append(sim.fields())
    for o in outs[1:]:
        for x, y in zip(outs[0], o):
            assert np.array_equal(bits(x), bits(y))


def test_mass_is_conserved_in_periodic_box():
    sim = HostSim(OracleBackend(19, SRT, FP32, 0), 16, 16, 16, nu=0.02)
    rho, u, flags = scenario(16, 16, 16, seed=2, solid_frac=0.0)
    load_scenario(sim, rho, u, flags)
    sim.run(0)
    m0 = float(np.sum(sim.fields()[0].astype(np.float64) - 1.0))
    sim.run(40)
    m1 = float(np.sum(sim.fields()[0].astype(np.float64) - 1.0))
    assert abs(m1 - m0) < 1e-4 * 16 ** 3 * 1e-2


def test_solid_cells_bounce_back_after_two_steps():
    # a single +x moving population next to a wall comes back reversed two steps later
    b = OracleBackend(19, SRT, FP32, 0)
    sim = HostSim(b, 8, 4, 4, w=1.0)  # w=1: pure relaxation to equilibrium keeps the picture simple
    flags = np.zeros((4, 4, 8), np.uint8); flags[:, :, 7] = TYPE_S; flags[:, :, 0] = TYPE_S
    rho = np.ones((4, 4, 8), np.float32); u = [np.zeros((4, 4, 8), np.float32) for _ in range(3)]
    u[0][:, :, 3] = 0.05
    load_scenario(sim, rho, u, flags)
    sim.run(30)
    r, ux, uy, uz, fl = sim.fields()
    fluid = fl == 0
    assert np.isfinite(r).all() and abs(float(np.sum(r[fluid].astype(np.float64))) - fluid.sum()) < 1e-3
    assert np.abs(ux[~fluid]).max() == 0.0  # walls keep u=0 (initialize zeroes solid velocity, src/kernel.cpp:1379)


def test_poiseuille_profile():
    R, umax, tau = 15, 0.1, 1.0
    nu = (tau - 0.5) / 3.0
    H = 2 * (R + 1)
    f = 4.0 * umax * 1.0 * nu / R ** 2          # src/units.hpp:113
    sim = HostSim(OracleBackend(19, SRT, FP32, VOLUME_FORCE), H, 2, H, nu=nu, fy=f)
    zz, yy, xx = np.meshgrid(np.arange(H), np.arange(2), np.arange(H), indexing="ij")
    r = np.sqrt((xx + 0.5 - 0.5 * H) ** 2 + (zz + 0.5 - 0.5 * H) ** 2)
    flags = np.where(r <= 0.5 * H - 1.0 + 0.5, 0, TYPE_S).astype(np.uint8)  # cylinder() predicate radius = min(Nx,Nz)/2-1 around center()
    rr = np.sqrt((xx - (0.5 * H - 0.5)) ** 2 + (zz - (0.5 * H - 0.5)) ** 2)
    flags = np.where(rr ** 2 <= (0.5 * H - 1.0) ** 2, 0, TYPE_S).astype(np.uint8)
    load_scenario(sim, np.ones((H, 2, H), np.float32), [np.zeros((H, 2, H), np.float32)] * 3, flags)
    sim.run(3000)
    _, ux, uy, uz, _ = sim.fields()
    unum = np.sqrt(ux.astype(np.float64) ** 2 + uy.astype(np.float64) ** 2 + uz.astype(np.float64) ** 2)[:, 1, :]
    rmid = r[:, 1, :]
    uref = umax * (R ** 2 - rmid ** 2) / R ** 2
    m = rmid < R
    err = np.sqrt(np.sum((unum[m] - uref[m]) ** 2) / np.sum(uref[m] ** 2))
    assert err < 0.06, err


def test_taylor_green_decay_rate():
    N, nu, A = 32, 0.02, 0.05
    sim = HostSim(OracleBackend(19, SRT, FP32, 0), N, N, 2, nu=nu)
    zz, yy, xx = np.meshgrid(np.arange(2), np.arange(N), np.arange(N), indexing="ij")
    k = 2 * np.pi / N
    fx, fy = xx + 0.5 - 0.5 * N, yy + 0.5 - 0.5 * N
    ux = (A * np.cos(k * fx) * np.sin(k * fy)).astype(np.float32)
    uy = (-A * np.sin(k * fx) * np.cos(k * fy)).astype(np.float32)
    rho = (1.0 - A * A * 3.0 / 4.0 * (np.cos(2 * k * fx) + np.cos(2 * k * fy))).astype(np.float32)  # src/setup.cpp:63-80
    load_scenario(sim, rho, [ux, uy, np.zeros_like(ux)], np.zeros((2, N, N), np.uint8))
    sim.run(0)
    e0 = float(np.sum(sim.fields()[1].astype(np.float64) ** 2))
    T = 200
    sim.run(T)
    e1 = float(np.sum(sim.fields()[1].astype(np.float64) ** 2))
    rate = -np.log(e1 / e0) / (2 * T)          # amplitude decays as exp(-nu*(kx^2+ky^2)*t)
    assert rate == pytest.approx(nu * 2 * k * k, rel=0.03)
