"""The CPU oracle against oracle/_ref (the reference's kernel.cpp compiled natively) on fresh random scenes.
Skipped where oracle/_ref is not built (it needs the mounted reference tree; the golden-vector tests cover that case)."""
import numpy as np
import pytest
from helpers import (OracleBackend, RefBackend, HostSim, scenario, load_scenario, ref_available,
                     FP32, FP16S, FP16C, SRT, TRT)

VARIANTS = [(19, SRT, FP32, 0), (19, SRT, FP16S, 0), (19, SRT, FP16C, 0), (19, TRT, FP32, 0), (19, SRT, FP32, 1), (19, SRT, FP32, 2),
            (19, TRT, FP16S, 3), (19, SRT, FP32, 4), (27, SRT, FP32, 0), (27, TRT, FP32, 3), (27, SRT, FP16S, 0), (27, TRT, FP16C, 3),
            (19, SRT, FP32, 8), (19, TRT, FP16S, 11), (27, SRT, FP16C, 8), (19, SRT, FP16S, 8),  # feature bit 3: SUBGRID
            (19, SRT, FP32, 16), (19, TRT, FP16S, 19), (27, SRT, FP16C, 18), (19, SRT, FP16S, 24)]  # feature bit 4: MOVING_BOUNDARIES


def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def run(cls, v, dims, D, steps, seed):
    Q, coll, st, feat = v
    f = (1e-4, -2e-4, 3e-4) if feat & 1 else (0.0, 0.0, 0.0)
    sim = HostSim(cls(Q, coll, st, feat), *dims, *D, nu=0.03, fx=f[0], fy=f[1], fz=f[2])
    rho, u, flags = scenario(sim.Nx, sim.Ny, sim.Nz, seed=seed, eq_frac=0.03 if feat & 2 else 0.0)
    load_scenario(sim, rho, u, flags)
    sim.run(steps)
    if feat & 16:  # MOVING_BOUNDARIES: the boundaries change speed (one stops, one starts), the marks are refreshed, the run continues
        for d in sim.dom:
            solid = (d.flags & 3) == 1
            idx = np.flatnonzero(solid)
            N = d.flags.size
            if idx.size > 1:
                for a in range(3): d.u[a * N + idx[0]] = 0.0
                d.u[idx[1]] = np.float32(0.02)
        sim._communicate("ruf")
        sim.update_moving_boundaries()
        sim.run(2)
    return list(sim.fields()) + [d.fi for d in sim.dom]


@pytest.mark.parametrize("v", VARIANTS, ids=[f"q{v[0]}c{v[1]}s{v[2]}f{v[3]}" for v in VARIANTS])
def test_oracle_equals_reference_device_code(v):
    if not ref_available(*v):
        pytest.skip("oracle/_ref variant not built")
    for dims, D, steps, seed in [((9, 7, 5), (1, 1, 1), 3, 1), ((12, 8, 6), (2, 1, 1), 4, 2), ((8, 8, 8), (2, 2, 2), 5, 3), ((6, 10, 8), (1, 2, 2), 2, 4)]:
        a = run(OracleBackend, v, dims, D, steps, seed)
        r = run(RefBackend, v, dims, D, steps, seed)
        for x, y in zip(a, r):
            assert np.array_equal(bits(x), bits(y)), (v, dims, D, steps)
