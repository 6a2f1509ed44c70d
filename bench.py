#!/usr/bin/env python3
"""bench.py -- headline benchmark of the LBM hot path: MLUPs/s (million lattice-cell updates per second) and fraction of
the HBM roofline, on the workload BASELINE.json's metric is quoted on.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

N=1 workload (BASELINE.json configs[1]): D3Q19 SRT, FP16S-compressed DDFs, 512^3 fully periodic box, default fields.
N>1 (torchrun, one process per GPU): weak scaling, every GPU owns one 512^3 FP16S domain of a Dx x Dy x Dz decomposition --
by default the reference's convention 2x1x1, 2x2x1, 2x2x2 (README.md:1425; x is decomposed first), `--split` for any other --
and exchanges halos with its neighbours over NVLink (CUDA-IPC peer memory; no host round trip).
A "step" is one LBM time step (stream_collide over every cell, plus the halo exchange when decomposed).
Before anything is timed, every run verifies the path it is about to measure: a 64^3-per-GPU perturbed-IC case of the same
lattice model runs 20 steps on the same devices / peer links and rank 0 compares every domain's rho and u with the CPU oracle's
decomposed run bit for bit ("verify" in the JSON line; a mismatch ends the run with a non-zero exit code).

One JSON line on stdout (rank 0). Keys follow the driver contract; see DESIGN.md section "Measurement".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (velocity_set, collision, storage, features, per-GPU box, description)
    "d3q19_srt_fp16s_512": (19, "srt", "fp16s", 0, (512, 512, 512), "D3Q19 SRT FP16S 512^3 periodic box (BASELINE configs[1])"),
    "d3q19_srt_fp32_256": (19, "srt", "fp32", 0, (256, 256, 256), "D3Q19 SRT FP32 256^3 periodic box (reference BENCHMARK setup, BASELINE configs[0])"),
    "d3q19_srt_fp32_512": (19, "srt", "fp32", 0, (512, 512, 512), "D3Q19 SRT FP32 512^3 periodic box"),
    "d3q19_srt_fp16c_512": (19, "srt", "fp16c", 0, (512, 512, 512), "D3Q19 SRT FP16C 512^3 periodic box"),
    "d3q19_srt_fp16s_256": (19, "srt", "fp16s", 0, (256, 256, 256), "D3Q19 SRT FP16S 256^3 periodic box"),
    "d3q19_srt_fp16c_1024": (19, "srt", "fp16c", 0, (1024, 1024, 1024), "D3Q19 SRT FP16C 1024^3 per GPU; with --gpus 8 --split 2,2,2 the 2048^3 domain of BASELINE configs[3] (run with --no-e2e: 18 GB of host fields per GPU)"),
    "d3q27_trt_fp32_windtunnel": (27, "trt", "fp32", 3, (256, 512, 256), "D3Q27 TRT FP32 wind tunnel with sphere, TYPE_E faces + VOLUME_FORCE (BASELINE configs[2], half size)"),
    "d3q27_trt_fp32_windtunnel_full": (27, "trt", "fp32", 3, (512, 1024, 512), "D3Q27 TRT FP32 wind tunnel with sphere, TYPE_E faces + VOLUME_FORCE, 512x1024x512 (BASELINE configs[2], SURVEY 8d C3)"),
    "d3q19_srt_fp32_256_cavity_mb": (19, "srt", "fp32", 16, (256, 256, 256), "D3Q19 SRT FP32 256^3 lid-driven cavity in the reference's own formulation (src/setup.cpp:212-249): six TYPE_S walls, the lid moves with u=(0,0.1,0), MOVING_BOUNDARIES, Re=1000"),
    "d3q19_srt_fp32_512_subgrid": (19, "srt", "fp32", 8, (512, 512, 512), "D3Q19 SRT FP32 512^3 periodic box with the SUBGRID (Smagorinsky-Lilly) model, SURVEY 8f rank 2"),
    "d3q19_srt_fp16s_512_subgrid": (19, "srt", "fp16s", 8, (512, 512, 512), "D3Q19 SRT FP16S 512^3 periodic box with the SUBGRID model"),
    "d3q19_srt_fp32_256_cavity": (19, "srt", "fp32", 2, (256, 256, 256), "D3Q19 SRT FP32 256^3 lid-driven cavity inside the hot-path feature set: TYPE_S walls, TYPE_E lid u=(0,0.1,0), Re=1000 (SURVEY 8d C1w)"),
}
DEFAULT_OVERLAP = False  # multi-GPU step: overlap the halo exchange with the interior cells
DEFAULT_WORKLOAD = "d3q19_srt_fp16s_512"
# weak scaling: every GPU keeps the full per-GPU box. Default domain grid = the reference's multi-GPU convention (README.md:1425,
# src/setup.cpp:20-23): x first. `--split 1,2,4` stacks along z and y instead (faces that are contiguous in memory).
SPLITS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def make_scene(workload, fx, Nx, Ny, Nz):
    """flags and u_y of the two scenes with boundaries (global arrays, shape (Nz,Ny,Nx))"""
    import numpy as np
    zz, yy, xx = np.meshgrid(np.arange(Nz), np.arange(Ny), np.arange(Nx), indexing="ij", sparse=True)
    flags = np.zeros((Nz, Ny, Nx), np.uint8)
    if "cavity_mb" in workload:  # the reference's cavity: all six faces solid, the lid (z = Nz-1) moves along +y
        for sl in [np.s_[0, :, :], np.s_[-1, :, :], np.s_[:, 0, :], np.s_[:, -1, :], np.s_[:, :, 0], np.s_[:, :, -1]]:
            flags[sl] = fx.TYPE_S
        uy = np.zeros((Nz, Ny, Nx), np.float32); uy[-1, :, :] = 0.1
        return flags, uy
    if "cavity" in workload:  # walls on five faces, equilibrium lid on z = Nz-1 moving along +y
        for sl in [np.s_[0, :, :], np.s_[:, 0, :], np.s_[:, -1, :], np.s_[:, :, 0], np.s_[:, :, -1]]:
            flags[sl] = fx.TYPE_S
        flags[-1, :, :] = fx.TYPE_E
        uy = np.zeros((Nz, Ny, Nx), np.float32); uy[-1, :, :] = 0.1
        return flags, uy
    # wind tunnel: sphere + TYPE_E faces
    flags[(xx - Nx / 2) ** 2 + (yy - Ny / 4) ** 2 + (zz - Nz / 2) ** 2 <= (Nx / 8) ** 2] = fx.TYPE_S
    for sl in [np.s_[0, :, :], np.s_[-1, :, :], np.s_[:, 0, :], np.s_[:, -1, :], np.s_[:, :, 0], np.s_[:, :, -1]]:
        flags[sl] = fx.TYPE_E
    return flags, np.where(flags == fx.TYPE_S, 0.0, 0.075).astype(np.float32)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clocks and throttle reasons with nvidia-smi while the timed region runs"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.samples)}


def reference_arm(args, wl_name):
    """--impl reference: the reference's own device code for this path (oracle/_ref: src/kernel.cpp compiled natively) on the
    host cores, all OpenMP threads, on a bounded sample of the workload. Falls back to the oracle port if _ref is absent."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import helpers as H
    Q, coll, st, feat, box, desc = WORKLOADS[wl_name]
    collision, storage = {"srt": H.SRT, "trt": H.TRT}[coll], {"fp32": H.FP32, "fp16s": H.FP16S, "fp16c": H.FP16C}[st]
    kind = "reference" if H.ref_available(Q, collision, storage, feat) else "port"
    cores = os.cpu_count() or 1  # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    backend = (H.RefBackend if kind == "reference" else H.OracleBackend)(Q, collision, storage, feat, threads=cores)
    n = 128 if cores < 32 else 192  # bounded sample: an n^3 sub-box of the workload, same kernels, same IC
    sim = H.HostSim(backend, n, n, n, nu=1.0)
    sim.run(0)
    for _ in range(max(1, args.warmup // 4)):
        sim.run(1)
    t0 = time.perf_counter()
    sim.run(args.steps)
    dt = time.perf_counter() - t0
    mlups = n ** 3 * args.steps / dt * 1e-6
    sample = f"{n}^3 sub-box of the workload ({desc}), {args.steps} steps, {cores} OpenMP threads"
    line = {"metric": "MLUPs/s", "value": round(mlups, 2), "unit": "MLUPs/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dt / args.steps * 1e3, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": st, "data": "synthetic",
            "config": {"workload": wl_name, "description": desc, "sample": sample},
            "cpu_baseline": {"value": round(mlups, 2), "unit": "MLUPs/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(mlups, 2), "unit": "MLUPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(wl_name, budget_s=15.0):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    Q, coll, st, feat, box, desc = WORKLOADS[wl_name]
    collision, storage = {"srt": H.SRT, "trt": H.TRT}[coll], {"fp32": H.FP32, "fp16s": H.FP16S, "fp16c": H.FP16C}[st]
    kind = "reference" if H.ref_available(Q, collision, storage, feat) else "port"
    cores = os.cpu_count() or 1
    backend = (H.RefBackend if kind == "reference" else H.OracleBackend)(Q, collision, storage, feat, threads=cores)
    n = 128
    sim = H.HostSim(backend, n, n, n, nu=1.0)
    sim.run(2)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s and steps < 200:
        sim.run(2); steps += 2
    dt = time.perf_counter() - t0
    return {"value": round(n ** 3 * steps / dt * 1e-6, 2), "unit": "MLUPs/s", "cores": cores, "kind": kind,
            "sample": f"{n}^3 sub-box of the workload, {steps} steps, {cores} OpenMP threads, "
                      + ("reference kernel.cpp device code compiled natively (oracle/_ref)" if kind == "reference" else "oracle port")}


def verify(fx, comm, Q, collision, storage, feat, split, device, overlap):
    """the path about to be timed, checked first: 64^3 cells per GPU, perturbed initial condition (plus random TYPE_S / TYPE_E cells where
    the workload has boundaries), 20 steps on the same devices and peer links; rank 0 compares every domain with the oracle bit for bit"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import helpers as H
    Dx, Dy, Dz = split
    n, steps, nu = 64, 20, 0.05
    Nx, Ny, Nz = n * Dx, n * Dy, n * Dz
    f = (1e-5, -2e-5, 3e-5) if feat & 1 else (0.0, 0.0, 0.0)
    rho, u, flags = H.scenario(Nx, Ny, Nz, seed=21, solid_frac=0.05 if feat & (2 | 16) else 0.0, eq_frac=0.03 if feat & 2 else 0.0)
    sim = fx.LBM(Nx, Ny, Nz, nu, *f, Dx=Dx, Dy=Dy, Dz=Dz, velocity_set=Q, collision=collision, storage=storage, features=feat,
                 comm=comm, devices=None if comm else [device], overlap=overlap)
    sim.rho.set_global(rho); [sim.u.set_global(u[a], a) for a in range(3)]; sim.flags.set_global(flags)
    sim.run(steps)
    sim.rho.read_from_device(); sim.u.read_from_device()
    mine = [sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2)]  # zero outside the domain this process owns
    (d, _), = sim.local_domains()
    sim.close()
    parts = comm.allgather((d, mine)) if comm is not None else [(d, mine)]
    rank = comm.rank if comm is not None else 0
    result = None
    if rank == 0:
        ref = H.HostSim(H.OracleBackend(Q, collision, storage, feat, threads=os.cpu_count() or 1), Nx, Ny, Nz, Dx, Dy, Dz, nu=nu, fx=f[0], fy=f[1], fz=f[2])
        H.load_scenario(ref, rho, u, flags)
        ref.run(steps)
        want = ref.fields()[:4]
        bad = 0
        for dd, arrs in parts:
            x, y, z = (dd % (Dx * Dy)) % Dx, (dd % (Dx * Dy)) // Dx, dd // (Dx * Dy)
            sl = (slice(z * n, (z + 1) * n), slice(y * n, (y + 1) * n), slice(x * n, (x + 1) * n))
            bad += sum(int(np.sum(a[sl].view(np.uint32) != b[sl].view(np.uint32))) for a, b in zip(arrs, want))
        result = {"ok": bad == 0, "split": [Dx, Dy, Dz], "grid": [Nx, Ny, Nz], "steps": steps, "mismatching_values": bad,
                  "what": "rho and u of every domain vs the CPU oracle's decomposed run, bit for bit, on the devices and peer links of the timed run"}
    if comm is not None:
        result = comm.allgather(result)[0]
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--overlap", type=int, default=-1, help="multi-GPU: 1 = shell/exchange on one stream, interior on another; 0 = one stream; -1 = library default")
    ap.add_argument("--reserve", type=int, default=-1, help="resident-block slots the interior kernel leaves free for the exchange (with --overlap 1)")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the workload's box is the GLOBAL grid, divided among the GPUs")
    ap.add_argument("--split", default="", help="Dx,Dy,Dz domain grid for multi-GPU runs (default: z first, then y)")
    ap.add_argument("--variant", type=int, default=0, help="kernel choice (fx3d_set_kernel_variant): 0 library default (bulk-copy kernels where eligible), 1 general one-cell-per-thread, "
                    "2 or 4 vector kernel with that many cells per thread, 8 persistent kernel with the cp.async ring, 16 bulk copies wherever eligible (incl. bulk stores on row segments)")
    ap.add_argument("--no-verify", action="store_true", help="skip the bit-exact check against the oracle that precedes the timed region")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, args.workload)
        return 0

    import numpy as np
    import fluidx3d_b200 as fx
    from fluidx3d_b200 import capi, lbm as lbm_mod
    lbm_mod.VERBOSE = False
    lib = capi.lib()
    lib.set_kernel_variant(args.variant)
    Q, coll, st, feat, box, desc = WORKLOADS[args.workload]
    collision, storage = {"srt": fx.SRT, "trt": fx.TRT}[coll], {"fp32": fx.FP32, "fp16s": fx.FP16S, "fp16c": fx.FP16C}[st]
    n_gpus = args.gpus
    comm, dist, torch = None, None, None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        comm = fx.TorchComm()
        n_gpus = world
    if n_gpus not in SPLITS:
        raise SystemExit(f"--gpus must be one of {sorted(SPLITS)}")
    Dx, Dy, Dz = SPLITS[n_gpus] if world > 1 else (1, 1, 1)
    if args.split:
        Dx, Dy, Dz = (int(v) for v in args.split.split(","))
        if Dx * Dy * Dz != n_gpus: raise SystemExit("--split must multiply to the number of GPUs")
    if world == 1 and n_gpus > 1:
        raise SystemExit("launch multi-GPU runs with torchrun (one process per GPU)")
    Nx, Ny, Nz = (box[0], box[1], box[2]) if args.strong else (box[0] * Dx, box[1] * Dy, box[2] * Dz)
    device = local_rank if world > 1 else 0
    K, W = args.steps, args.warmup

    def barrier():
        if dist is not None:
            dist.barrier()

    overlap = DEFAULT_OVERLAP if args.overlap < 0 else bool(args.overlap)
    if args.reserve >= 0: lib.set_interior_reserve(args.reserve)
    verified = None
    if not args.no_verify:
        verified = verify(fx, comm, Q, collision, storage, feat, (Dx, Dy, Dz), device, overlap)
        if not verified["ok"]:
            if rank == 0: print(json.dumps({"metric": "MLUPs/s", "value": None, "verify": verified, "error": "the CUDA path differs from the oracle; nothing was timed"}), flush=True)
            return 3
    # ---------------- device-resident arm: `value` ----------------
    force = (0.0, 1e-6, 0.0) if feat & 1 else (0.0, 0.0, 0.0)
    nu = 0.1 * (Nz - 2) / 1000.0 if "cavity" in args.workload else 1.0
    sim = fx.LBM(Nx, Ny, Nz, nu, *force, Dx=Dx, Dy=Dy, Dz=Dz, velocity_set=Q, collision=collision, storage=storage, features=feat,
                 comm=comm, devices=None if comm else [device], host_fields=bool(feat & (2 | 16)), benchmark=True, overlap=overlap)
    scene_flags, scene_uy = None, None
    if feat & (2 | 16):
        scene_flags, scene_uy = make_scene(args.workload, fx, Nx, Ny, Nz)
        sim.flags.set_global(scene_flags); sim.u.set_global(scene_uy, 1)
    (d0, dom), = sim.local_domains()
    sim.run(0)
    sim.run(W, sync=True)
    ev0, ev1 = C.c_void_p(), C.c_void_p()
    lib.event_create(dom.device, C.byref(ev0)); lib.event_create(dom.device, C.byref(ev1))
    sampler = ClockSampler(dom.device)
    launches0 = lib.launches()
    kinds0 = lib.kernel_kind_counts()
    barrier(); lib.stream_sync(dom.device, dom.stream)
    sampler.start()
    lib.event_record(dom.device, ev0, dom.stream)
    sim.run(K, sync=False)
    lib.event_record(dom.device, ev1, dom.stream)
    lib.event_sync(dom.device, ev1)
    barrier()
    ms = C.c_float(0.0)
    lib.event_elapsed_ms(ev0, ev1, C.byref(ms))
    sim.finish()
    clocks = sampler.summary()
    launches = lib.launches() - launches0
    kinds = [b - a for a, b in zip(kinds0, lib.kernel_kind_counts())]
    kernel_name = lib.KERNEL_KINDS[kinds.index(max(kinds))]  # the stream_collide form that actually ran in the timed region
    ms_total = ms.value
    if dist is not None:  # max over ranks of the device time
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    cells = Nx * Ny * Nz
    mlups = cells * K / (ms_total * 1e-3) * 1e-6
    # algorithmic bytes per cell and step: 2*Q*sizeof(fpxx)+1 (+16 with UPDATE_FIELDS), SURVEY 8d. The reference's own accounting
    # (bandwidth_bytes_per_cell_device, src/lbm.cpp:52-64) adds Q-1 neighbour-flag bytes for MOVING_BOUNDARIES builds; this path reads
    # those flags only for the few TYPE_MS cells, so they are not counted here (that would flatter the roofline fraction)
    bytes_per_cell = 2 * Q * (4 if st == "fp32" else 2) + 1 + (16 if feat & 4 else 0)
    peak, peak_src = measured_peaks()
    per_gpu_gbs = (cells / n_gpus) * bytes_per_cell * K / (ms_total * 1e-3) * 1e-9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": round(per_gpu_gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(per_gpu_gbs / peak, 4),
                "traffic": traffic, "peak_source": peak_src, "kernel": kernel_name,
                "bytes_per_cell_per_step": bytes_per_cell, "cells_per_launch": cells // n_gpus}
    if roofline["frac"] > 1.0:  # the driver's figure is a torch copy_ (6.45 TB/s on this pool); the FP32 kernel moves its bytes faster than that copy
        roofline["note"] = f"achieved exceeds the measured copy bandwidth; {per_gpu_gbs / 8000.0:.3f} of the 8 TB/s data-sheet HBM3e figure"
    sim.close()

    # ---------------- end-to-end arm through the host API with host buffers: `e2e` ----------------
    e2e = None
    if not args.no_e2e:
        sim = fx.LBM(Nx, Ny, Nz, nu, *force, Dx=Dx, Dy=Dy, Dz=Dz, velocity_set=Q, collision=collision, storage=storage, features=feat,
                     comm=comm, devices=None if comm else [device], host_fields=True, benchmark=True, overlap=overlap)
        if feat & (2 | 16):
            sim.flags.set_global(scene_flags); sim.u.set_global(scene_uy, 1)
        (d0, dom), = sim.local_domains()
        h2d = dom.rho.nbytes + dom.u.nbytes + dom.flags.nbytes
        d2h = dom.rho.nbytes + dom.u.nbytes
        sim.run(0); sim.run(3)  # warm
        sim.reset()
        barrier()
        t0 = time.perf_counter()
        sim.run(K)                       # initialize(): H2D of rho,u,flags from pinned host memory + initialize kernel, then K steps
        sim.rho.read_from_device()       # update_fields kernel + D2H
        sim.u.read_from_device()
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        e2e = {"value": round(cells * K / dt * 1e-6, 1), "unit": "MLUPs/s", "h2d_bytes_per_step": int(h2d * n_gpus / K), "d2h_bytes_per_step": int(d2h * n_gpus / K),
               "what": f"LBM ctor fields in pinned host memory -> run({K}) (H2D rho,u,flags + initialize + {K} steps) -> rho/u.read_from_device() (update_fields + D2H); wall clock"}
        sim.close()

    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload)

    if rank == 0:
        published = {"d3q19_srt_fp32_256": 42152.0, "d3q19_srt_fp16s_256": 55609.0}  # README.md:1147, 1x B200, 256^3
        line = {"metric": "MLUPs/s", "value": round(mlups, 1), "unit": "MLUPs/s", "n_gpus": n_gpus, "steps": K, "warmup": W,
                "ms_per_step": round(ms_total / K, 5), "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
                "vs_baseline": round(mlups / published[args.workload], 4) if (args.workload in published and n_gpus == 1) else None,
                "dtype": "f32 arithmetic, " + st + " DDF storage", "data": "synthetic",
                "config": {"workload": args.workload, "description": desc, "global_grid": [Nx, Ny, Nz], "domains": [Dx, Dy, Dz], "overlap": bool(overlap and n_gpus > 1),
                           "cache": "DDF working set per GPU (%.1f GB) is far larger than the 126 MB L2; no flush needed" % (cells / n_gpus * Q * (4 if st == "fp32" else 2) * 1e-9)},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "verify": verified}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
