"""Multi-GPU correctness check (run under torchrun, one process per GPU): a decomposed run with direct NVLink peer
exchange over CUDA IPC must be bit-identical to the CPU oracle's decomposed run (and hence to the single-domain run)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
import fluidx3d_b200 as fx
from fluidx3d_b200 import lbm as lbm_mod
import helpers as H
lbm_mod.VERBOSE = False
SPLITS = {2: [(2, 1, 1), (1, 1, 2)], 4: [(2, 2, 1), (1, 4, 1)], 8: [(2, 2, 2), (4, 1, 2)]}
ok_all = True
for D in SPLITS[world]:
  for overlap in (False, True):
    for (Q, coll, st, feat) in [(19, 0, 0, 0), (19, 0, 2, 0), (27, 1, 1, 3)]:
        dims, steps = (64, 32, 32), 9
        f = (1e-4, 0.0, -1e-4) if feat & 1 else (0.0, 0.0, 0.0)
        comm = fx.TorchComm()
        sim = fx.LBM(*dims, 0.05, *f, Dx=D[0], Dy=D[1], Dz=D[2], velocity_set=Q, collision=coll, storage=st, features=feat, comm=comm, overlap=overlap)
        rho, u, flags = H.scenario(*dims, seed=12, eq_frac=0.03 if feat & 2 else 0.0)
        sim.rho.set_global(rho); [sim.u.set_global(u[a], a) for a in range(3)]; sim.flags.set_global(flags)
        sim.run(steps)
        for m in (sim.rho, sim.u, sim.flags): m.read_from_device()
        mine = [sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2)]
        parts = [None] * world
        dist.all_gather_object(parts, [a.tobytes() for a in mine])
        if rank == 0:
            tot = [np.zeros(a.size, np.uint32) for a in mine]
            for p in parts:
                for k in range(4): tot[k] |= np.frombuffer(p[k], np.uint32)
            ref = H.HostSim(H.OracleBackend(Q, coll, st, feat), *dims, *D, nu=0.05, fx=f[0], fy=f[1], fz=f[2])
            H.load_scenario(ref, rho, u, flags); ref.run(steps)
            ok = all(np.array_equal(t, w.view(np.uint32).ravel()) for t, w in zip(tot, ref.fields()[:4]))
            ok_all &= ok
            print(f"dist_check world={world} D={D} overlap={int(overlap)} Q={Q} coll={coll} storage={st} feat={feat}: {'OK' if ok else 'MISMATCH'}", flush=True)
        sim.close()
        dist.barrier()
if rank == 0:
    print("DIST_CHECK", "PASS" if ok_all else "FAIL", flush=True)
dist.destroy_process_group()
