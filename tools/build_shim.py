#!/usr/bin/env python3
"""Build the reference WITH the B200 shim: `shim/_build/<cfg>/athena_<pgen>`.

The binary is the reference (main.cpp, ParameterInput, Mesh / MeshBlock, its C++ problem
generators, Mesh::Initialize, the polling task scheduler, outputs ...) compiled from the
sources where they lie under /root/reference by the recipe of oracle/build_ref.py (its own
configure.py / Makefile are not run, no reference source is copied or modified), except that two
translation units are left out and replaced by this repository's shim/*.cpp:

    src/task_list/time_integrator.cpp  ->  shim/b200_time_integrator.cpp  (task bodies = C ABI)
    src/hydro/new_blockdt.cpp          ->  shim/b200_new_blockdt.cpp      (device dt + upload)
                                       +   shim/b200_bridge.cpp

and it links libathena_b200.so.  This is what a maintainer's `configure.py -b200` would produce
(INTEGRATION.md); tests/test_shim_cpu.py runs it against the emulated library on the CPU,
tests/test_gpu_shim.py against the real one on the GPU, both bit-for-bit against the goldens
that the UNMODIFIED reference binary produced.

Usage:  python tools/build_shim.py [--ref /root/reference] [cfg[:pgen,...] ...]
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref  # noqa: E402  (the reference build recipe; building, not using, the checker)

SHIM = os.path.join(ROOT, "shim")
OUT = os.path.join(SHIM, "_build")
LIBDIR = os.path.join(ROOT, "athena-gamma_b200")
REPLACED = ("task_list__time_integrator.o", "hydro__new_blockdt.o")
SHIM_SRC = ("b200_bridge.cpp", "b200_time_integrator.cpp", "b200_new_blockdt.cpp")

# what the tests need: the five BASELINE configurations + user hooks, scalars, isothermal
DEFAULT = {
    "mhd_hlld_ng2": ["blast", "linear_wave", "shk_cloud", "usersrc", "shock_tube"],
    "mhd_hlld_ng3": ["orszag_tang", "blast"],
    "hydro_hllc_ng3": ["kh"],
    "hydro_hllc_ng2": ["shock_tube", "blast"],
    "mhd_hlld_ng2_s1": ["kh"],
    "hydro_hlle_iso_ng2": ["blast"],
}


def exe_path(cfg, pgen):
    return os.path.join(OUT, cfg, "athena_" + pgen)


def have(cfg, pgen):
    return os.path.exists(exe_path(cfg, pgen))


def build(cfg, pgens, ref, jobs):
    src = os.path.join(ref, "src")
    refroot = os.path.join(build_ref.OUT, cfg)
    obj = os.path.join(refroot, "obj")
    inc = os.path.join(refroot, "inc")
    # the reference objects of this configuration (compiled by the oracle recipe; no-op if fresh)
    build_ref.build(cfg, ref, jobs)
    out = os.path.join(OUT, cfg)
    os.makedirs(out, exist_ok=True)
    incs = ["-I", src, "-I", inc, "-I", os.path.join(inc, "sub"), "-I", os.path.join(ROOT, "include"),
            "-I", SHIM]
    shim_objs = []
    for s in SHIM_SRC:
        o = os.path.join(out, s[:-4] + ".o")
        cmd = ["g++"] + build_ref.CXXFLAGS + incs + ["-c", os.path.join(SHIM, s), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit("shim compile failed: %s\n%s" % (s, r.stderr[-4000:]))
        shim_objs.append(o)
    mhd, flux, ng, all_pgens, nscalars, eos = build_ref.cfg_tuple(cfg)
    common = []
    for s in build_ref.source_list(src, mhd, flux, eos):
        rel = os.path.relpath(s, src).replace("/", "__")[:-4] + ".o"
        if rel in REPLACED:
            continue
        common.append(os.path.join(obj, rel))
    for p in pgens:
        pobj = os.path.join(obj, "pgen__" + p + ".o")
        if not os.path.exists(pobj):
            raise SystemExit("pgen %s is not part of oracle/build_ref.py's %s" % (p, cfg))
        exe = exe_path(cfg, p)
        cmd = (["g++"] + build_ref.CXXFLAGS + ["-s", "-o", exe] + common + shim_objs + [pobj]
               + ["-L", LIBDIR, "-lathena_b200", "-Wl,--enable-new-dtags",
                  "-Wl,-rpath,$ORIGIN/../../../athena-gamma_b200"])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit("shim link failed: %s\n%s" % (exe, r.stderr[-4000:]))
        print("built", exe)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("cfgs", nargs="*")
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(a.ref, "src")):
        print("reference tree not present at %s: keeping prebuilt shim/_build as is" % a.ref)
        return 0
    if not os.path.exists(os.path.join(LIBDIR, "libathena_b200.so")):
        raise SystemExit("build athena-gamma_b200/libathena_b200.so first (__graft_entry__.build)")
    want = {}
    for c in a.cfgs:
        cfg, _, pg = c.partition(":")
        want[cfg] = pg.split(",") if pg else DEFAULT.get(cfg) or build_ref.cfg_tuple(cfg)[3]
    for cfg, pgens in (want or DEFAULT).items():
        build(cfg, [p.split(":")[-1] for p in pgens], a.ref, a.jobs)
    return 0


if __name__ == "__main__":
    sys.exit(main())
