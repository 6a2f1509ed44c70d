"""Where does a multi-GPU step spend its time? (torchrun, one process per GPU). Times, per rank, with CUDA events:
collide only / collide + rendezvous / exchange only / full step; also the host time needed to enqueue them."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
import fluidx3d_b200 as fx
from fluidx3d_b200 import lbm as lbm_mod, capi
lbm_mod.VERBOSE = False
lib = capi.lib()
split = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,1,2").split(","))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
K = 50
sim = fx.LBM(n * split[0], n * split[1], n * split[2], 1.0, Dx=split[0], Dy=split[1], Dz=split[2], velocity_set=19, storage=fx.FP16S, comm=fx.TorchComm(), host_fields=False, benchmark=True)
(d0, dom), = sim.local_domains()
sim.run(5)
ev0, ev1 = C.c_void_p(), C.c_void_p()
lib.event_create(dom.device, C.byref(ev0)); lib.event_create(dom.device, C.byref(ev1))
def timed(name, body):
    dist.barrier(); lib.stream_sync(dom.device, dom.stream)
    lib.event_record(dom.device, ev0, dom.stream)
    t0 = time.perf_counter()
    for _ in range(K): body()
    host = (time.perf_counter() - t0) / K * 1e3
    lib.event_record(dom.device, ev1, dom.stream); lib.event_sync(dom.device, ev1)
    ms = C.c_float(0.0); lib.event_elapsed_ms(ev0, ev1, C.byref(ms))
    sim.finish()
    print(f"rank {rank} {name:28s} device {ms.value / K:8.3f} ms/step   host enqueue {host:8.3f} ms/step", flush=True)
    dist.barrier()
def collide(): dom.enqueue_stream_collide()
def collide_rdv(): dom.enqueue_stream_collide(); sim._barrier(None)
def rdv(): sim._barrier(None)
def exch():
    for axis, Dn in enumerate(split):
        if Dn > 1:
            p, m = sim._peers[sim._neighbour(d0, axis, +1)], sim._peers[sim._neighbour(d0, axis, -1)]
            lib.exchange_fi(C.byref(dom.lat), axis, dom.t, p["fi"], m["fi"], dom.stream)
def full(): sim.do_time_step()
def fused():
    if sim._fused: lib.stream_collide_fused(C.byref(dom.lat), dom.t, dom.fx, dom.fy, dom.fz, sim._fused[d0], dom.stream)
def xfaces():
    if sim.Dx > 1: sim._communicate("fi", axes=(0,))
for name, body in [("collide only", collide), ("rendezvous only", rdv), ("collide + rendezvous", collide_rdv), ("exchange only (direct pull)", exch), ("fused collide only", fused), ("x faces only (staged)", xfaces), ("full step", full), ("collide only (again)", collide)]:
    timed(name, body)
sim.close()
dist.destroy_process_group()
