"""athena-gamma_b200: B200-native (sm_100a CUDA) per-MeshBlock hydro/MHD update of the Athena++
fork pabolmasov/Athena-gamma, behind the reference's own surface.

  csrc/            CUDA kernels + host runtime + C ABI  -> libathena_b200.so
  lib.py           ctypes binding of include/athena_b200.h
  athinput.py      ParameterInput (athinput `<block> key = value` files, cmd-line overrides)
  mesh.py          Mesh / MeshBlock host mirror driving the C ABI
  pgen/            problem generators (the ProblemGenerator hook) for the five configs
  build.py         nvcc build of the extension (sm_100a, -fmad=false)
"""
from .athinput import ParameterInput  # noqa: F401
from .mesh import Mesh, MeshPlan  # noqa: F401
from . import lib, pgen, build  # noqa: F401

__all__ = ["ParameterInput", "Mesh", "MeshPlan", "lib", "pgen", "build"]
