import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import fluidx3d_b200 as fx
from fluidx3d_b200 import capi, lbm as lbm_mod
import helpers as H
lbm_mod.VERBOSE = False
lib = capi.lib()
bad = C.c_uint64(0); lib.selftest_division(0, 1 << 28, C.byref(bad)); print("division selftest mismatches:", bad.value)
def run(v, dims, steps, variant):
    Q, coll, st, feat = v
    lib.set_kernel_variant(variant)
    sim = fx.LBM(*dims, 0.05, velocity_set=Q, collision=coll, storage=st, features=feat)
    rho, u, flags = H.scenario(*dims, seed=3)
    sim.rho.set_global(rho); [sim.u.set_global(u[a], a) for a in range(3)]; sim.flags.set_global(flags)
    sim.run(steps)
    for m in (sim.rho, sim.u, sim.flags): m.read_from_device()
    out = [sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2)]
    sim.close()
    ref = H.HostSim(H.OracleBackend(Q, coll, st, feat), *dims, nu=0.05)
    H.load_scenario(ref, rho, u, flags); ref.run(steps)
    want = ref.fields()[:4]
    return out, want, flags
for v in [(19,0,0,0),(19,0,1,0),(19,0,2,0)]:
    for variant in (0,1):
        for steps in (1,2):
            out, want, flags = run(v, (32,12,10), steps, variant)
            for name, a, b in zip(("rho","ux","uy","uz"), out, want):
                d = a.view(np.uint32) != b.view(np.uint32)
                if d.any():
                    idx = np.argwhere(d)
                    z,y,x = idx[0]
                    print(v, "variant", variant, "steps", steps, name, "mismatches", int(d.sum()), "of", d.size, "first at zyx", idx[0], "got %r want %r" % (a[z,y,x], b[z,y,x]), "ulps", int(a.view(np.int32)[z,y,x]) - int(b.view(np.int32)[z,y,x]), "flag", flags[z,y,x], "x%4 =", x%4)
                else:
                    print(v, "variant", variant, "steps", steps, name, "OK")
