"""The CPU oracle against the committed golden vectors (generated from the reference's own device code, see
oracle/make_golden.py): bit-exact rho/u/flags and raw DDF buffers for every variant, decomposition and step parity."""
import glob
import os
import numpy as np
import pytest
from helpers import OracleBackend, HostSim, load_scenario, ROOT

FILES = sorted(f for f in glob.glob(os.path.join(ROOT, "tests", "golden", "q*.npz")))


def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def run_case(backend_cls, g):
    Q, coll, st, feat, Nx, Ny, Nz, Dx, Dy, Dz, steps = (int(v) for v in g["meta"])
    f = g["force"]
    sim = HostSim(backend_cls(Q, coll, st, feat), Nx, Ny, Nz, Dx, Dy, Dz, w=float(g["w"]), fx=float(f[0]), fy=float(f[1]), fz=float(f[2]))
    load_scenario(sim, g["in_rho"], list(g["in_u"]), g["in_flags"])
    sim.run(steps)
    rho, ux, uy, uz, flags = sim.fields()
    return rho, np.stack([ux, uy, uz]), flags, np.stack([d.fi for d in sim.dom])


def test_golden_files_present():
    assert len(FILES) >= 12


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    rho, u, flags, fi = run_case(OracleBackend, g)
    assert np.array_equal(flags, g["out_flags"])
    assert np.array_equal(bits(rho), bits(g["out_rho"]))
    assert np.array_equal(bits(u), bits(g["out_u"]))
    assert np.array_equal(bits(fi), bits(g["out_fi"]))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_golden_w_is_the_decimal_round_trip(path):
    g = np.load(path)
    assert np.float32(OracleBackend().w_from_nu(float(g["nu"]))) == g["w"]
