"""The one-process-per-domain path (what bench.py runs under torchrun) on CPU: two / four processes over torch.distributed
(gloo), each owning one domain and mapping its neighbours' buffers through the IPC entry points, ordered by the rendezvous
counters. Runs against the host-emulation build of the library (tests/emul), so it checks the host logic -- handle exchange,
neighbour tables, rendezvous sequencing, exchange calls -- and the exchange kernels' addressing; result must equal the oracle."""
import os
import subprocess
import sys
import numpy as np
import pytest
from helpers import ROOT

WORKER = r'''
import os, sys, json
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch.distributed as dist
dist.init_process_group(backend="gloo")
from fluidx3d_b200 import capi, lbm as lbm_mod
from fluidx3d_b200.lbm import LBM, TorchComm
import helpers as H
lbm_mod.VERBOSE = False
lib = capi.Lib(os.path.join(ROOT, "tests", "_build", "libfx3d_emul.so"))
D = tuple(int(v) for v in os.environ["FX3D_TEST_D"].split(","))
Q, coll, st, feat = (int(v) for v in os.environ["FX3D_TEST_V"].split(","))
dims, steps = (16, 8, 8), 5
comm = TorchComm()
sim = LBM(*dims, 0.05, Dx=D[0], Dy=D[1], Dz=D[2], velocity_set=Q, collision=coll, storage=st, features=feat, comm=comm, lib=lib)
rho, u, flags = H.scenario(*dims, seed=6)
sim.rho.set_global(rho); [sim.u.set_global(u[a], a) for a in range(3)]; sim.flags.set_global(flags)
sim.run(steps)
for m in (sim.rho, sim.u, sim.flags): m.read_from_device()
mine = [sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2)]
parts = [None]*comm.world_size
dist.all_gather_object(parts, [a.tobytes() for a in mine])
if comm.rank == 0:
    # each rank filled only its own block of the global arrays (zeros elsewhere; blocks are disjoint) -> sum the bit patterns
    tot = [np.zeros(a.size, np.uint32) for a in mine]
    for p in parts:
        for k in range(4): tot[k] |= np.frombuffer(p[k], np.uint32)
    ref = H.HostSim(H.OracleBackend(Q, coll, st, feat), *dims, *D, nu=0.05)
    H.load_scenario(ref, rho, u, flags); ref.run(steps)
    want = ref.fields()[:4]
    ok = all(np.array_equal(t, w.view(np.uint32).ravel()) for t, w in zip(tot, want))
    print("RESULT", "OK" if ok else "MISMATCH")
sim.close()
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("D,v", [((2, 1, 1), (19, 0, 0, 0)), ((1, 2, 1), (19, 0, 1, 0)), ((2, 2, 1), (27, 1, 2, 0))], ids=["2x1x1-fp32", "1x2x1-fp16s", "2x2x1-q27trt16c"])
def test_one_process_per_domain_matches_oracle(D, v, tmp_path):
    subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "tests", "emul")], check=True, capture_output=True)
    script = tmp_path / "worker.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + WORKER)
    n = D[0] * D[1] * D[2]
    env = dict(os.environ, FX3D_TEST_D=",".join(map(str, D)), FX3D_TEST_V=",".join(map(str, v)), OMP_NUM_THREADS="1")
    port = 29500 + (os.getpid() + n * 7 + v[2]) % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "RESULT OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
