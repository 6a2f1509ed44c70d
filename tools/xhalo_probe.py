"""Single GPU, two domains side by side: time stream_collide alone on a domain with a halo along `axis` (0 x, 1 y, 2 z)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fluidx3d_b200 as fx
from fluidx3d_b200 import lbm as lbm_mod, capi
lbm_mod.VERBOSE = False
lib = capi.lib()
axis = int(sys.argv[1]) if len(sys.argv) > 1 else 0
st = {"fp16s": fx.FP16S, "fp32": fx.FP32, "fp16c": fx.FP16C}[sys.argv[2] if len(sys.argv) > 2 else "fp16s"]
K = int(sys.argv[3]) if len(sys.argv) > 3 else 30
n = 512
D = [1, 1, 1]; D[axis] = 2
sim = fx.LBM(n * D[0], n * D[1], n * D[2], 1.0, Dx=D[0], Dy=D[1], Dz=D[2], velocity_set=19, storage=st, devices=[0, 0], host_fields=False, benchmark=True)
sim.run(2)
d0, dom = sim.local_domains()[0]
ev0, ev1 = C.c_void_p(), C.c_void_p()
lib.event_create(dom.device, C.byref(ev0)); lib.event_create(dom.device, C.byref(ev1))
lib.stream_sync(dom.device, dom.stream)
lib.event_record(dom.device, ev0, dom.stream)
for _ in range(K): dom.enqueue_stream_collide()
lib.event_record(dom.device, ev1, dom.stream); lib.event_sync(dom.device, ev1)
ms = C.c_float(0.0); lib.event_elapsed_ms(ev0, ev1, C.byref(ms))
print(f"halo axis {axis}: domain {dom.Nx}x{dom.Ny}x{dom.Nz} collide {ms.value / K:.3f} ms/step -> {n**3 / (ms.value / K) * 1e-3:.0f} MLUPs/s", flush=True)
sim.close()
