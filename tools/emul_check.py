#!/usr/bin/env python3
"""Check a tuning configuration of the whole-row kernel on the CPU before spending GPU time: builds the emulation library (tests/emul) with extra
defines into a scratch directory and compares a set of whole-row cases (one-row and many-row tiles, x halos, fused y/z delivery) with the oracle.
usage: emul_check.py <tag> [-DFX3D_ROW_K_32=2 ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import OracleBackend, HostSim, scenario, load_scenario, FP32, FP16S, FP16C, SRT, TRT
from fluidx3d_b200 import capi
from fluidx3d_b200 import lbm as lbm_mod
from fluidx3d_b200.lbm import LBM
lbm_mod.VERBOSE = False
tag, extra = sys.argv[1], " ".join(sys.argv[2:])
out = f"/tmp/fx3d_emul_{tag}"
subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "tests", "emul"), f"OUT={out}", f"EXTRA={extra}"], check=True, capture_output=True)
lib = capi.Lib(os.path.join(out, "libfx3d_emul.so"))
bits = lambda a: a.view(np.uint32) if a.dtype == np.float32 else a
CASES = [((19, SRT, FP16S, 0), (512, 6, 7), (1, 1, 1)), ((19, SRT, FP32, 0), (512, 7, 6), (1, 1, 1)), ((19, TRT, FP16C, 3), (256, 12, 6), (1, 1, 1)),
         ((27, TRT, FP32, 3), (64, 16, 6), (1, 1, 1)), ((19, SRT, FP32, 1), (1024, 6, 12), (2, 1, 2)), ((19, SRT, FP16S, 2), (1024, 12, 6), (2, 2, 1)),
         ((19, SRT, FP32, 0), (64, 32, 12), (1, 2, 2)), ((19, SRT, FP16S, 24), (64, 32, 12), (1, 2, 2)), ((27, SRT, FP16S, 0), (128, 8, 12), (1, 1, 2))]
bad = 0
for v, dims, D in CASES:
    f = (1e-4, -2e-4, 3e-4) if v[3] & 1 else (0.0, 0.0, 0.0)
    before = lib.kernel_kind_counts()
    for steps in (1, 2, 5):
        sim = LBM(*dims, 0.05, *f, Dx=D[0], Dy=D[1], Dz=D[2], velocity_set=v[0], collision=v[1], storage=v[2], features=v[3], lib=lib)
        rho, u, flags = scenario(sim.Nx, sim.Ny, sim.Nz, seed=3, eq_frac=0.03 if v[3] & 2 else 0.0)
        sim.rho.set_global(rho); [sim.u.set_global(u[a], a) for a in range(3)]; sim.flags.set_global(flags)
        sim.run(steps)
        for m in (sim.rho, sim.u, sim.flags): m.read_from_device()
        got = (sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2), sim.flags.get_global())
        sim.close()
        ref = HostSim(OracleBackend(*v), *dims, *D, nu=0.05, fx=f[0], fy=f[1], fz=f[2])
        load_scenario(ref, rho, u, flags)
        ref.run(steps)
        ok = all(np.array_equal(bits(a), bits(b)) for a, b in zip(got, ref.fields()))
        if not ok: bad += 1
    ran = [b - a for a, b in zip(before, lib.kernel_kind_counts())]
    print(v, dims, D, "OK" if ok else "MISMATCH", "row-kernel launches:", ran[3], "others:", sum(ran) - ran[3], flush=True)
print("emul_check", tag, "FAILED" if bad else "passed")
sys.exit(1 if bad else 0)
