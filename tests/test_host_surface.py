"""The C++ host surface (fluidx3d_b200/host: LBM, LBM_Domain, Memory<T>, Device, scenes) must compile for every supported
combination of the reference's defines.hpp switches, and must refuse the unsupported ones with a clear message. Compile-only
(g++ -fsyntax-only): no GPU, no linking."""
import os
import subprocess
import pytest
from helpers import ROOT

HOST = os.path.join(ROOT, "fluidx3d_b200", "host")
SRC = ["main.cpp", "lbm.cpp", "info.cpp", "shapes.cpp", "setup.cpp"]


def compile_only(defs):
    cmd = ["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-pthread", "-Wall", "-Wno-unused-parameter"] + defs + SRC
    return subprocess.run(cmd, cwd=HOST, capture_output=True, text=True)


GOOD = [
    [],                                                                                                       # defines.hpp as shipped: BENCHMARK, D3Q19 SRT FP16S
    ["-DFX3D_CUSTOM_DEFINES", "-DSCENE_TAYLOR_GREEN", "-DD3Q19", "-DSRT"],
    ["-DFX3D_CUSTOM_DEFINES", "-DSCENE_POISEUILLE", "-DD3Q19", "-DSRT", "-DVOLUME_FORCE"],
    ["-DFX3D_CUSTOM_DEFINES", "-DSCENE_CAVITY", "-DD3Q19", "-DSRT", "-DFP16S", "-DEQUILIBRIUM_BOUNDARIES"],
    ["-DFX3D_CUSTOM_DEFINES", "-DSCENE_CAVITY", "-DD3Q19", "-DSRT", "-DFP16C", "-DMOVING_BOUNDARIES"],
    ["-DFX3D_CUSTOM_DEFINES", "-DSCENE_WINDTUNNEL", "-DD3Q27", "-DTRT", "-DEQUILIBRIUM_BOUNDARIES", "-DVOLUME_FORCE", "-DSUBGRID", "-DUPDATE_FIELDS"],
]
BAD = [
    (["-DFX3D_CUSTOM_DEFINES", "-DSCENE_TAYLOR_GREEN", "-DD3Q19", "-DSRT", "-DSURFACE"], "not part of the B200 hot-path build"),
    (["-DFX3D_CUSTOM_DEFINES", "-DSCENE_TAYLOR_GREEN", "-DD3Q19", "-DSRT", "-DINTERACTIVE_GRAPHICS"], "graphics are not part"),
    (["-DFX3D_CUSTOM_DEFINES", "-DSCENE_TAYLOR_GREEN", "-DD2Q9", "-DSRT"], "only D3Q19 and D3Q27"),
]


@pytest.mark.parametrize("defs", GOOD, ids=lambda d: "+".join(x[2:] for x in d) or "shipped")
def test_host_compiles(defs):
    r = compile_only(defs)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.parametrize("defs,message", BAD, ids=lambda v: "+".join(x[2:] for x in v) if isinstance(v, list) else None)
def test_host_refuses_unsupported_extensions(defs, message):
    r = compile_only(defs)
    assert r.returncode != 0 and message in r.stderr
