"""Build libathena_b200.so in-tree with nvcc for sm_100a.

-fmad=false is load-bearing: the reference's default build has no FMA contraction
(configure.py:454, x86-64 SSE2), and bit-identical dt sequences require the same roundings.
"""
import fcntl
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in ("ab_kernels.cu", "ab_flux_nu.cu", "ab_mesh.cu",
                                               "ab_smr.cpp", "ab_smr_kernels.cu")]
HDR = [os.path.join(HERE, "csrc", f) for f in ("ab_kernels.h", "ab_types.h", "ab_physics.cuh",
                                               "ab_flux.cuh", "ab_batch.cuh", "ab_smr_cells.cuh", "ab_smr_exec.h")] + \
      [os.path.join(os.path.dirname(HERE), "include", "athena_b200.h")]
SO = os.path.join(HERE, "libathena_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "--threads", "0"]


STAMP = SO + ".srchash"


def source_hash():
    """sha256 over the sources, headers and flags the library is built from (mtimes do not
    survive the copy to a GPU box; contents do)"""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SRC + HDR:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


KERNEL_FILES = ("ab_kernels.cu", "ab_flux_nu.cu", "ab_types.h", "ab_physics.cuh", "ab_flux.cuh",
                "ab_batch.cuh")


def kernel_hash():
    """sha256 over the sources of the kernels on the cycle path and the compiler flags: what an
    ncu capture of those kernels depends on (the host glue in ab_mesh.cu / ab_smr.cpp, the launcher
    declarations and the refinement kernels may change without invalidating it)"""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for name in KERNEL_FILES:
        h.update(name.encode())
        with open(os.path.join(HERE, "csrc", name), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(SO) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != source_hash()


def build(force=False, verbose=False):
    """Compile unless the library on disk was built from exactly these sources.  Safe to call
    from several ranks at once: one builds, the others wait on the lock file."""
    if not force and not needs_build():
        return SO
    with open(SO + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():      # another rank built it while we waited
            return SO
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        tmp = SO + ".tmp%d" % os.getpid()
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SRC + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        os.replace(tmp, SO)
        with open(STAMP, "w") as fh:
            fh.write(source_hash() + "\n")
        if verbose:
            print(r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
