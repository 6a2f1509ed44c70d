#!/usr/bin/env python3
"""Generate tests/golden/*.npz from oracle/_ref -- the reference's own kernel.cpp device code compiled natively
(oracle/ref/build_ref.py) and driven through the reference's host sequencing (tests/helpers.HostSim restates
src/lbm.cpp:881-953,1343-1390). Runs only in the container where /root/reference is mounted; the vectors are
committed so that the oracle and the CUDA path can be checked anywhere. TEST INFRASTRUCTURE ONLY.

Each file holds the inputs (rho,u,flags on the global grid, nu/w, force), the step count, and the outputs:
rho/u/flags as lbm.rho/u/flags.read_from_device() would return them, plus every domain's raw DDF buffer."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from helpers import *  # noqa

OUT = os.path.join(ROOT, "tests", "golden")
CASES = [  # (Q, collision, storage, features, (Nx,Ny,Nz), (Dx,Dy,Dz), steps, nu, force)
    (19, SRT, FP32, 0, (12, 10, 8), (1, 1, 1), 7, 0.05, (0, 0, 0)),
    (19, SRT, FP32, 0, (12, 10, 8), (2, 1, 2), 6, 0.05, (0, 0, 0)),
    (19, SRT, FP16S, 0, (12, 10, 8), (1, 1, 1), 7, 0.05, (0, 0, 0)),
    (19, SRT, FP16C, 0, (12, 10, 8), (1, 2, 1), 7, 0.05, (0, 0, 0)),
    (19, TRT, FP32, 0, (12, 10, 8), (1, 1, 1), 5, 0.02, (0, 0, 0)),
    (19, SRT, FP32, VOLUME_FORCE, (12, 10, 8), (1, 1, 1), 5, 1.0 / 6.0, (1e-4, -2e-4, 3e-4)),
    (19, SRT, FP32, EQUILIBRIUM_BOUNDARIES, (12, 10, 8), (1, 1, 1), 5, 0.05, (0, 0, 0)),
    (19, TRT, FP16S, 3, (12, 10, 8), (2, 2, 2), 6, 0.03, (2e-4, 0, -1e-4)),
    (27, SRT, FP32, 0, (12, 10, 8), (1, 1, 1), 5, 0.05, (0, 0, 0)),
    (27, TRT, FP32, 3, (12, 10, 8), (2, 2, 2), 6, 0.01, (0, 1e-4, 0)),
    (27, SRT, FP16S, 0, (12, 10, 8), (1, 1, 2), 5, 0.05, (0, 0, 0)),
    (27, TRT, FP16C, 3, (12, 10, 8), (1, 1, 1), 5, 0.04, (1e-4, 1e-4, 1e-4)),
    (19, SRT, FP32, 8, (12, 10, 8), (1, 1, 1), 6, 0.002, (0, 0, 0)),               # SUBGRID (feature bit 3), low viscosity so that the eddy term matters
    (19, TRT, FP16S, 11, (12, 10, 8), (2, 1, 2), 6, 0.002, (2e-4, 0, -1e-4)),
    (27, SRT, FP16C, 8, (12, 10, 8), (1, 1, 1), 5, 0.004, (0, 0, 0)),
    (19, SRT, FP32, 16, (12, 10, 8), (1, 1, 1), 6, 0.05, (0, 0, 0)),               # MOVING_BOUNDARIES (feature bit 4): the random scene's solids keep their velocity
    (19, TRT, FP16S, 19, (12, 10, 8), (2, 1, 2), 6, 0.03, (2e-4, 0, -1e-4)),
    (27, SRT, FP16C, 18, (12, 10, 8), (1, 1, 1), 5, 0.05, (0, 0, 0)),
]


def case_name(Q, coll, st, feat, dims, D, steps):
    return f"{ref_variant_name(Q, coll, st, feat)}_{dims[0]}x{dims[1]}x{dims[2]}_d{D[0]}{D[1]}{D[2]}_t{steps}"


def main():
    os.makedirs(OUT, exist_ok=True)
    for (Q, coll, st, feat, dims, D, steps, nu, f) in CASES:
        b = RefBackend(Q, coll, st, feat)
        sim = HostSim(b, *dims, *D, nu=nu, fx=f[0], fy=f[1], fz=f[2])
        rho, u, flags = scenario(*dims, seed=11, eq_frac=0.04 if feat & EQUILIBRIUM_BOUNDARIES else 0.0)
        load_scenario(sim, rho, u, flags)
        sim.run(steps)
        o_rho, o_ux, o_uy, o_uz, o_flags = sim.fields()
        np.savez_compressed(os.path.join(OUT, case_name(Q, coll, st, feat, dims, D, steps) + ".npz"),
                            meta=np.array([Q, coll, st, feat, *dims, *D, steps], dtype=np.int64), nu=np.float32(nu), w=np.float32(sim.w),
                            force=np.array(f, dtype=np.float32), in_rho=rho, in_u=np.stack(u), in_flags=flags,
                            out_rho=o_rho, out_u=np.stack([o_ux, o_uy, o_uz]), out_flags=o_flags,
                            out_fi=np.stack([d.fi for d in sim.dom]))
        print("wrote", case_name(Q, coll, st, feat, dims, D, steps))
    # FORCE_FIELD: the scenario, the sequence and the outputs of tests/test_force_field.run_host, from the reference's device code
    import test_force_field as T
    for v, dims, D, steps, seed in [((19, SRT, FP32, 33), (12, 10, 8), (1, 1, 1), 5, 21), ((19, TRT, FP16S, 35), (12, 10, 8), (2, 1, 2), 4, 22),
                                    ((27, SRT, FP16C, 33), (12, 10, 8), (1, 2, 1), 4, 23), ((19, SRT, FP32, 34), (12, 10, 8), (1, 1, 1), 5, 24)]:
        _, out = T.run_host(RefBackend, v, dims, D, steps, seed)
        name = f"ff_{ref_variant_name(*v)}_{dims[0]}x{dims[1]}x{dims[2]}_d{D[0]}{D[1]}{D[2]}_t{steps}"
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=np.array([*v, *dims, *D, steps, seed], dtype=np.int64), **{f"out{k}": np.asarray(x) for k, x in enumerate(out)})
        print("wrote", name)
    # storage codec known-answer vectors from the reference's device converters (src/kernel.cpp:848-859)
    b = RefBackend(19, SRT, FP16C, 0)
    codes = np.arange(65536, dtype=np.uint32)
    dec = np.array([b.lib.ref_half_to_float_custom(int(c)) for c in codes], dtype=np.float32)
    rng = np.random.default_rng(5)
    xs = np.concatenate([dec, np.nextafter(dec, np.float32(4)), np.nextafter(dec, np.float32(-4)),
                         rng.uniform(-2.5, 2.5, 30000).astype(np.float32),
                         (rng.uniform(-1, 1, 30000) * 10.0 ** rng.uniform(-12, 0, 30000)).astype(np.float32),
                         np.array([0.0, -0.0, 2.0, -2.0, 3.999, 1e-38, 1e-45, 6.1e-5, 2.98e-8, 1.49e-8, 1.4901161e-8], dtype=np.float32)]).astype(np.float32)
    enc = np.array([b.lib.ref_float_to_half_custom(float(x)) for x in xs], dtype=np.uint16)
    np.savez_compressed(os.path.join(OUT, "fp16c_codec.npz"), decode_all_codes=dec.view(np.uint32), encode_in=xs.view(np.uint32), encode_out=enc)
    print("wrote fp16c_codec")


if __name__ == "__main__":
    main()
