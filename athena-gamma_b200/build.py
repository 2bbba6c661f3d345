"""Build libathena_b200.so in-tree with nvcc for sm_100a.

-fmad=false is load-bearing: the reference's default build has no FMA contraction
(configure.py:454, x86-64 SSE2), and bit-identical dt sequences require the same roundings.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in ("ab_kernels.cu", "ab_flux_nu.cu", "ab_mesh.cu",
                                               "ab_smr.cpp", "ab_smr_kernels.cu")]
HDR = [os.path.join(HERE, "csrc", f) for f in ("ab_kernels.h", "ab_types.h", "ab_physics.cuh",
                                               "ab_flux.cuh", "ab_smr_cells.cuh", "ab_smr_exec.h")] + \
      [os.path.join(os.path.dirname(HERE), "include", "athena_b200.h")]
SO = os.path.join(HERE, "libathena_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "--threads", "0"]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(f) > t for f in SRC + HDR)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + SRC + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
