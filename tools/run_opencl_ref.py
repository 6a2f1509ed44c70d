#!/usr/bin/env python3
"""Run the reference's own OpenCL build (oracle/_ref/opencl, built by oracle/ref/build_opencl_ref.py from the untouched
/root/reference tree) beside this repository's CUDA path on the same B200 (SURVEY 8d "Reference beside it (i)").

  * its BENCHMARK setup (src/setup.cpp:5-36) for FP32 / FP16S / FP16C at 256^3 -> "Peak MLUPs/s"
  * the same box sizes as bench.py's workloads through tests/scenes/file_scene.cpp (host-clock MLUPs/s over N steps)
  * parity: identical perturbed initial conditions, N steps, fields dumped by read_from_device(): flags must be identical,
    rho/u within 1e-5 (FP32) / 1e-3 (FP16 storage) relative -- the north_star's literal correctness clause

The NVIDIA driver on the box ships libnvidia-opencl.so.1 but no /etc/OpenCL/vendors/*.icd; the ICD loader of the CUDA toolkit
(libOpenCL.so.1) is pointed at it with OCL_ICD_FILENAMES. Writes profiles/r02_reference_opencl.json (and gpurun_out/ copy).
"""
import glob, json, os, re, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
BIN = os.path.join(ROOT, "oracle", "_ref", "opencl")
OUTDIR = os.path.join(ROOT, "gpurun_out")


def opencl_env():
    env = dict(os.environ)
    icd = [p for pat in ("/usr/lib/libnvidia-opencl.so.1", "/usr/lib/x86_64-linux-gnu/libnvidia-opencl.so.1", "/usr/lib64/libnvidia-opencl.so.1") for p in glob.glob(pat)]
    if icd:
        env["OCL_ICD_FILENAMES"] = icd[0]
    return env, (icd[0] if icd else None)


def run_ref(binary, env_extra, timeout=600):
    env, _ = opencl_env()
    env.update({k: str(v) for k, v in env_extra.items()})
    t0 = time.time()
    r = subprocess.run([os.path.join(BIN, binary)], env=env, capture_output=True, text=True, timeout=timeout, cwd=OUTDIR)
    return r.returncode, r.stdout.replace("\r", "\n"), r.stderr, time.time() - t0


def main():
    import numpy as np
    import helpers as H
    import fluidx3d_b200 as fx
    from fluidx3d_b200 import lbm as lbm_mod
    lbm_mod.VERBOSE = False
    os.makedirs(OUTDIR, exist_ok=True)
    report = {"icd": opencl_env()[1], "benchmark_256": {}, "bench_same_sizes": {}, "parity": []}
    log = open(os.path.join(OUTDIR, "opencl_ref.log"), "w")
    quick = "--quick" in sys.argv

    # ---- 1. the reference's own BENCHMARK setup ----
    for st in ("fp32", "fp16s", "fp16c"):
        rc, out, err, dt = run_ref(f"FluidX3D_bench_{st}", {})
        log.write(f"==== bench_{st} rc={rc} {dt:.1f}s\n{out[-3000:]}\n{err[-2000:]}\n")
        m = re.search(r"Peak MLUPs/s = (\d+)", out)
        dev = re.search(r"\| Device Name\s+\| (.*?)\s+\|", out)
        report["benchmark_256"][st] = {"peak_mlups": int(m.group(1)) if m else None, "rc": rc, "seconds": round(dt, 1), "device": dev.group(1) if dev else None}
        print("reference BENCHMARK", st, report["benchmark_256"][st], flush=True)
        if rc != 0 and not m:
            print(out[-1500:], err[-1500:])
            break

    # ---- 2. same sizes as bench.py's workloads ----
    if not quick:
        for name, binary, n, extra in (("d3q19_srt_fp32_512", "FluidX3D_q19_srt_fp32_f0", (512, 512, 512), {}), ("d3q19_srt_fp16s_512", "FluidX3D_q19_srt_fp16s_f0", (512, 512, 512), {}),
                                       ("d3q19_srt_fp16c_512", "FluidX3D_q19_srt_fp16c_f0", (512, 512, 512), {})):
            rc, out, err, dt = run_ref(binary, {"FX3D_REF_N": "%d,%d,%d" % n, "FX3D_REF_STEPS": 20, "FX3D_REF_BENCH": 200, **extra})
            log.write(f"==== {name} rc={rc} {dt:.1f}s\n{out[-2000:]}\n{err[-2000:]}\n")
            m = re.search(r"FX3D_REF_RESULT mlups=([\d.]+)", out)
            report["bench_same_sizes"][name] = {"mlups": float(m.group(1)) if m else None, "rc": rc, "steps": 200, "timing": "host clock around lbm.run(200) (one finish_queue per step, as the reference runs)"}
            print("reference", name, report["bench_same_sizes"][name], flush=True)

    # ---- 3. parity on identical inputs ----
    cases = [  # variant, dims, D, steps, nu, force, scenario kwargs
        ("q19_srt_fp32_f0", (64, 64, 64), (1, 1, 1), 100, 0.05, None, {}),
        ("q19_srt_fp16s_f0", (96, 64, 48), (1, 1, 1), 100, 0.05, None, {}),
        ("q19_srt_fp16c_f0", (64, 64, 64), (1, 1, 1), 100, 0.05, None, {}),
        ("q19_srt_fp32_f0", (64, 64, 64), (2, 2, 2), 20, 0.05, None, {}),
        ("q19_srt_fp16c_f0", (64, 64, 64), (2, 1, 1), 20, 0.05, None, {}),
        ("q27_trt_fp32_f3", (64, 128, 64), (1, 1, 1), 50, 0.02, (0.0, 1e-6, 0.0), {"eq_frac": 0.03}),
        ("q19_trt_fp16s_f3", (64, 64, 64), (1, 2, 1), 50, 0.03, (1e-5, 0.0, 2e-5), {"eq_frac": 0.03}),
        ("q27_srt_fp16c_f0", (48, 48, 48), (1, 1, 1), 50, 0.1, None, {}),
        ("q19_srt_fp32_f16", (64, 64, 64), (1, 1, 1), 50, 0.05, None, {}),
        ("q19_srt_fp16s_f8", (64, 64, 64), (1, 1, 1), 50, 0.01, None, {}),
        ("q19_srt_fp32_f0", (64, 64, 64), (1, 1, 1), 100, 0.02, None, {"scene": "taylor_green"}),
        ("q19_srt_fp16s_f0", (96, 64, 48), (1, 1, 1), 100, 0.02, None, {"scene": "taylor_green"}),
        ("q19_srt_fp16c_f0", (64, 64, 64), (2, 2, 2), 100, 0.02, None, {"scene": "taylor_green"}),
        ("q27_trt_fp32_f3", (64, 128, 64), (1, 1, 1), 100, 0.01, (0.0, 1e-6, 0.0), {"scene": "taylor_green"}),
        ("q19_srt_fp32_f0", (256, 256, 256), (1, 1, 1), 20, 0.02, None, {"scene": "taylor_green"}),
    ]
    if quick: cases = cases[:3]
    if "--steps-sweep" in sys.argv: cases = [("q19_srt_fp32_f0", (64, 64, 64), (1, 1, 1), n, 0.05, None, {}) for n in (1, 2, 10, 50, 100)]
    stn = {"fp32": fx.FP32, "fp16s": fx.FP16S, "fp16c": fx.FP16C}
    for variant, dims, D, steps, nu, force, kw in cases:
        q, coll, st, f = variant.split("_")
        Q, feat = int(q[1:]), int(f[1:])
        Nx, Ny, Nz = dims
        kw = dict(kw)
        scene = kw.pop("scene", "noise")
        if scene == "noise":  # perturbed rho/u + 8 % random solid cells (tests/helpers.py: scenario)
            rho, u, flags = H.scenario(Nx, Ny, Nz, seed=11, **kw)
        else:  # Taylor-Green vortices (the initial condition of src/setup.cpp:50-61) around a solid sphere: the amplitude survives the run
            zz, yy, xx = np.meshgrid(np.arange(Nz), np.arange(Ny), np.arange(Nx), indexing="ij")
            a, b, c = 2 * np.pi * xx / Nx, 2 * np.pi * yy / Ny, 2 * np.pi * zz / Nz
            u = [(0.1 * np.cos(a) * np.sin(b) * np.sin(c)).astype(np.float32), (-0.1 * np.sin(a) * np.cos(b) * np.sin(c)).astype(np.float32), np.zeros((Nz, Ny, Nx), np.float32)]
            rho = (1.0 - 0.01 * 3.0 / 4.0 * (np.cos(2 * a) + np.cos(2 * b))).astype(np.float32)
            flags = np.zeros((Nz, Ny, Nx), np.uint8)
            flags[(xx - Nx / 2) ** 2 + (yy - Ny / 4) ** 2 + (zz - Nz / 2) ** 2 <= (Nx / 8) ** 2] = 1
            if feat & 2: flags[:, 0, :] = 2; flags[:, -1, :] = 2
        fin, fout = os.path.join(OUTDIR, "ref_in.bin"), os.path.join(OUTDIR, "ref_out.bin")
        with open(fin, "wb") as fh:
            np.array([Nx, Ny, Nz], np.uint32).tofile(fh); rho.tofile(fh); [u[a].tofile(fh) for a in range(3)]; flags.tofile(fh)
        env = {"FX3D_REF_IN": fin, "FX3D_REF_OUT": fout, "FX3D_REF_STEPS": steps, "FX3D_REF_NU": repr(nu), "FX3D_REF_D": "%d,%d,%d" % D}
        if force: env["FX3D_REF_F"] = ",".join(repr(v) for v in force)
        if os.path.exists(fout): os.remove(fout)
        rc, out, err, dt = run_ref("FluidX3D_" + variant, env)
        log.write(f"==== parity {variant} {dims} D={D} rc={rc} {dt:.1f}s\n{out[-1500:]}\n{err[-1500:]}\n")
        entry = {"variant": variant, "grid": dims, "domains": D, "steps": steps, "rc": rc}
        if rc != 0 or not os.path.exists(fout):
            entry["error"] = (out[-400:] + err[-400:]); report["parity"].append(entry); print(entry, flush=True); continue
        N = Nx * Ny * Nz
        raw = np.fromfile(fout, np.uint8)
        fl_ref = raw[16 * N:17 * N].reshape(Nz, Ny, Nx)
        f_ref = raw[:16 * N].view(np.float32).reshape(4, Nz, Ny, Nx)
        # ours, same inputs, D domains on this one device (the reference does the same: "Using single fastest device for all domains")
        sim = fx.LBM(Nx, Ny, Nz, nu, *(force or (0.0, 0.0, 0.0)), Dx=D[0], Dy=D[1], Dz=D[2], velocity_set=Q, collision=fx.SRT if coll == "srt" else fx.TRT,
                     storage=stn[st], features=feat, devices=[0] * (D[0] * D[1] * D[2]))
        sim.rho.set_global(rho); [sim.u.set_global(u[a], a) for a in range(3)]; sim.flags.set_global(flags)
        sim.run(steps)
        sim.rho.read_from_device(); sim.u.read_from_device(); sim.flags.read_from_device()
        ours = np.stack([sim.rho.get_global(), sim.u.get_global(0), sim.u.get_global(1), sim.u.get_global(2)])
        fl_ours = sim.flags.get_global()
        sim.close()
        fluid = (fl_ref & 3) != 1  # initialize() zeroes u of solid cells; they are not part of the comparison
        tol = 1e-5 if st == "fp32" else 1e-3
        # "relative" = relative to the scale of the field: rho ~ 1, u ~ the velocity amplitude the scene starts with (the random-noise
        # scenes decay by orders of magnitude within 100 steps, the Taylor-Green scenes keep their amplitude: u_scale_final is reported too)
        d_rho = float(np.max(np.abs(ours[0] - f_ref[0])[fluid] / np.abs(f_ref[0])[fluid]))
        u_abs = float(np.max(np.abs(ours[1:] - f_ref[1:])[:, fluid]))
        u0 = float(max(np.max(np.abs(u[a])[fluid]) for a in range(3)))
        u1 = float(np.max(np.abs(f_ref[1:])[:, fluid]))
        identical = float(np.mean((ours.view(np.uint32) == f_ref.view(np.uint32))[:, fluid]))
        entry.update({"scene": scene, "flags_identical": bool(np.array_equal(fl_ref, fl_ours)), "rho_max_rel": d_rho, "u_max_abs": u_abs, "u_scale_initial": u0, "u_scale_final": u1,
                      "u_max_rel": u_abs / u0, "bit_identical_fraction": round(identical, 6), "tolerance": tol,
                      "ok": bool(np.array_equal(fl_ref, fl_ours) and d_rho <= tol and u_abs / u0 <= tol), "ref_seconds": round(dt, 1)})
        report["parity"].append(entry)
        print(entry, flush=True)
        if "--diagnose" in sys.argv:  # where is the largest velocity difference, and what do the three implementations hold there?
            k = int(np.argmax(np.where(fluid[None], np.abs(ours[1:] - f_ref[1:]), 0)))
            a, z, y, x = np.unravel_index(k, ours[1:].shape)
            print("   worst u component", a, "at", (x, y, z), "ours", ours[1 + a, z, y, x], "ref", f_ref[1 + a, z, y, x], "rho ours/ref", ours[0, z, y, x], f_ref[0, z, y, x],
                  "flags around", flags[max(z - 1, 0):z + 2, max(y - 1, 0):y + 2, max(x - 1, 0):x + 2].ravel().tolist())
            diff = np.where(fluid[None], np.abs(ours[1:] - f_ref[1:]), 0)
            print("   cells with |du| > 1e-6:", int(np.sum(diff.max(axis=0) > 1e-6)), "of", int(fluid.sum()), "fluid cells; median |du|", float(np.median(diff.max(axis=0)[fluid])))
            np.savez_compressed(os.path.join(OUTDIR, f"ocl_case_{variant}_{steps}.npz"), ours=ours, ref=f_ref, flags=flags)
    log.close()
    for f in ("ref_in.bin", "ref_out.bin"):  # up to 285 MB each: not for the 64 MiB that travels back
        if os.path.exists(os.path.join(OUTDIR, f)): os.remove(os.path.join(OUTDIR, f))
    report["all_ok"] = all(e.get("ok") for e in report["parity"])
    json.dump(report, open(os.path.join(OUTDIR, "r02_reference_opencl.json"), "w"), indent=1)
    print("ALL PARITY CASES OK" if report["all_ok"] else "SOME PARITY CASES FAILED")
    return 0


if __name__ == "__main__":
    sys.exit(main())
