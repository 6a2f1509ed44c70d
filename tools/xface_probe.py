"""Single GPU, two x-decomposed domains side by side: a few full time steps, for an ncu launch list of the exchange kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fluidx3d_b200 as fx
from fluidx3d_b200 import lbm as lbm_mod
lbm_mod.VERBOSE = False
st = {"fp16s": fx.FP16S, "fp32": fx.FP32, "fp16c": fx.FP16C}[sys.argv[1] if len(sys.argv) > 1 else "fp16s"]
n = 512
sim = fx.LBM(2 * n, n, n, 1.0, Dx=2, Dy=1, Dz=1, velocity_set=19, storage=st, devices=[0, 0], host_fields=False, benchmark=True)
sim.run(6)
sim.close()
print("done")
