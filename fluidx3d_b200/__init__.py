"""fluidx3d_b200 -- B200-native (sm_100a CUDA) implementation of FluidX3D's lattice-Boltzmann hot path behind the
reference's host API. The compute path is libfx3d_cuda.so (C ABI in include/fx3d.h); this package holds its ctypes
binding (capi) and the Python mirror of the LBM host classes (lbm). The C++ host surface is in fluidx3d_b200/host/."""
from .capi import (FP32, FP16S, FP16C, SRT, TRT, VOLUME_FORCE, EQUILIBRIUM_BOUNDARIES, UPDATE_FIELDS, SUBGRID, MOVING_BOUNDARIES, FORCE_FIELD, TYPE_S, TYPE_E, Fx3dError)  # noqa: F401
from .lbm import LBM, LBM_Domain, Memory_Container, TorchComm  # noqa: F401
from .mesh import Mesh, read_stl  # noqa: F401
