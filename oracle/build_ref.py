#!/usr/bin/env python3
"""Build recipe for `oracle/_ref/`: the UNMODIFIED reference, compiled where it lies.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (athena-gamma_b200/) may use what
this script builds; it exists so that (1) the C restatement in oracle/athena_oracle.c can be
pinned bit-for-bit against the real reference, (2) tests can generate golden state dumps,
and (3) bench.py's `--impl reference` / `cpu_baseline` arm can time the reference's own CPU
implementation on the GPU box's host cores.

What it does (our own recipe -- the reference's configure.py / Makefile are NOT run):
  * instantiates the macro template `src/defs.hpp.in` into `oracle/_ref/<cfg>/inc/defs.hpp`
    (plain `@KEY@` substitution, the same keys configure.py:355-446 fills);
  * compiles each needed reference source file straight from /root/reference/src with
    `g++ -O3 -std=c++11 -fopenmp` (the reference's default flag set, configure.py:454,668;
    never -ffast-math / -march=native, which would break bitwise parity);
  * links one binary per problem generator: `oracle/_ref/<cfg>/athena_<pgen>`.
No reference source is copied into the repo; outputs go only under oracle/_ref/ (git-ignored,
not gpurun-ignored, so the binaries travel to the GPU box).

Usage:  python oracle/build_ref.py [--ref /root/reference] [--jobs 8] [cfg ...]
        cfg names: see CONFIGS below; default = all.
"""
import argparse
import concurrent.futures as cf
import glob
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

# cfg name -> (mhd?, flux file stem, nghost, [pgens][, {"nscalars": n, "eos": "isothermal"}])
CONFIGS = {
    "hydro_hllc_ng2": (False, "hllc", 2, ["shock_tube", "linear_wave", "blast", "kh",
                                          "shk_cloud"]),
    "hydro_hlle_ng2": (False, "hlle", 2, ["shock_tube", "linear_wave"]),
    "hydro_roe_ng2": (False, "roe", 2, ["shock_tube", "linear_wave"]),
    "hydro_hllc_ng3": (False, "hllc", 3, ["kh", "shock_tube", "linear_wave"]),
    # even NGHOST for PPM with mesh refinement (MeshRefinement ctor rejects odd NGHOST)
    "hydro_hllc_ng4": (False, "hllc", 4, ["kh", "blast"]),
    "hydro_hllc_ng4_s2": (False, "hllc", 4, ["kh"], {"nscalars": 2}),
    "mhd_hlld_ng2": (True, "hlld", 2, ["linear_wave", "blast", "orszag_tang", "shock_tube",
                                       "shk_cloud", "local:usersrc"]),
    "mhd_hlle_ng2": (True, "hlle", 2, ["linear_wave", "shock_tube"]),
    "mhd_roe_ng2": (True, "roe", 2, ["linear_wave", "shock_tube"]),
    "mhd_hlld_ng3": (True, "hlld", 3, ["orszag_tang", "linear_wave", "blast"]),
    # the fork's production solvers (confignotes: --flux=lhllc / --flux=lhlld)
    "hydro_lhllc_ng2": (False, "lhllc", 2, ["blast", "shock_tube", "kh"]),
    "mhd_lhlld_ng2": (True, "lhlld", 2, ["blast", "orszag_tang", "linear_wave"]),
    # LLF (rsolvers/hydro/llf.cpp, mhd/llf_mhd.cpp)
    "hydro_llf_ng2": (False, "llf", 2, ["blast"]),
    "mhd_llf_ng2": (True, "llf", 2, ["blast"]),
    "mhd_llf_iso_ng2": (True, "llf", 2, ["blast"], {"eos": "isothermal"}),
    # passive scalars (confignotes: --nscalars=1 with lhllc) and the isothermal EOS
    # (confignotes: --flux=hlle --eos=isothermal)
    "hydro_lhllc_ng2_s1": (False, "lhllc", 2, ["kh", "shock_tube", "local:usersrc"],
                           {"nscalars": 1}),
    "hydro_hllc_ng3_s2": (False, "hllc", 3, ["kh"], {"nscalars": 2}),
    "mhd_hlld_ng2_s1": (True, "hlld", 2, ["kh"], {"nscalars": 1}),
    "hydro_hlle_iso_ng2": (False, "hlle", 2, ["linear_wave", "blast", "kh", "local:usersrc"],
                           {"eos": "isothermal"}),
    "hydro_hlle_iso_ng2_s1": (False, "hlle", 2, ["kh"], {"eos": "isothermal", "nscalars": 1}),
    "mhd_hlld_iso_ng2": (True, "hlld", 2, ["linear_wave", "orszag_tang", "blast"],
                         {"eos": "isothermal"}),
    "mhd_hlle_iso_ng2": (True, "hlle", 2, ["linear_wave", "orszag_tang"], {"eos": "isothermal"}),
    # Roe's solver with the isothermal EOS (the NON_BAROTROPIC_EOS == 0 branches of roe*.cpp)
    "hydro_roe_iso_ng2": (False, "roe", 2, ["blast", "kh"], {"eos": "isothermal"}),
    "mhd_roe_iso_ng2": (True, "roe", 2, ["blast", "orszag_tang"], {"eos": "isothermal"}),
}


def cfg_tuple(cfg):
    t = CONFIGS[cfg]
    opt = t[4] if len(t) > 4 else {}
    return t[0], t[1], t[2], t[3], int(opt.get("nscalars", 0)), opt.get("eos", "adiabatic")

CXXFLAGS = ["-O3", "-std=c++11", "-fopenmp"]


def defs_for(mhd, flux, nghost, nscalars=0, eos="adiabatic"):
    iso = eos == "isothermal"
    d = {
        "PROBLEM": "oracle_multi",
        "COORDINATE_SYSTEM": "cartesian",
        "RSOLVER": flux,
        "EQUATION_OF_STATE": eos,
        "GENERAL_EOS": "0",
        "EOS_TABLE_ENABLED": "0",
        "NON_BAROTROPIC_EOS": "0" if iso else "1",
        "MAGNETIC_FIELDS_ENABLED": "1" if mhd else "0",
        "STS_ENABLED": "0",
        "SELF_GRAVITY_ENABLED": "0",
        "RELATIVISTIC_DYNAMICS": "0",
        "GENERAL_RELATIVITY": "0",
        "FRAME_TRANSFORMATIONS": "0",
        "SINGLE_PRECISION_ENABLED": "0",
        "H5_DOUBLE_PRECISION_ENABLED": "0",
        "FFT_OPTION": "NO_FFT",
        "MPI_OPTION": "NOT_MPI_PARALLEL",
        "OPENMP_OPTION": "OPENMP_PARALLEL",
        "HDF5_OPTION": "NO_HDF5OUTPUT",
        "DEBUG_OPTION": "NOT_DEBUG",
        "EXCEPTION_HANDLING_OPTION": "ENABLE_EXCEPTIONS",
        "COMPILER_CHOICE": "g++",
        "COMPILER_COMMAND": "g++",
        "COMPILER_FLAGS": " ".join(CXXFLAGS),
        # configure.py:370-423
        "NHYDRO_VARIABLES": "4" if iso else "5",
        "NFIELD_VARIABLES": "3" if mhd else "0",
        "NWAVE_VALUE": ("6" if iso else "7") if mhd else ("4" if iso else "5"),
        "NUMBER_PASSIVE_SCALARS": str(nscalars),
        "NUMBER_GHOST_CELLS": str(nghost),
    }
    return d


def source_list(src, mhd, flux, eos="adiabatic"):
    """Same selection rule as the reference's Makefile.in:27-58 (one EOS, one solver)."""
    pats = ["*.cpp", "bvals/*.cpp", "bvals/cc/*.cpp", "bvals/cc/fft_grav/*.cpp",
            "bvals/cc/hydro/*.cpp", "bvals/cc/mg/*.cpp", "bvals/fc/*.cpp",
            "bvals/orbital/*.cpp", "bvals/utils/*.cpp", "coordinates/*.cpp", "fft/*.cpp",
            "field/*.cpp", "field/field_diffusion/*.cpp", "gravity/*.cpp", "hydro/*.cpp",
            "hydro/srcterms/*.cpp", "hydro/hydro_diffusion/*.cpp", "inputs/*.cpp",
            "mesh/*.cpp", "multigrid/*.cpp", "orbital_advection/*.cpp", "outputs/*.cpp",
            "reconstruct/*.cpp", "scalars/*.cpp", "task_list/*.cpp", "utils/*.cpp"]
    files = []
    for p in pats:
        files += sorted(glob.glob(os.path.join(src, p)))
    eosf = eos + ("_mhd.cpp" if mhd else "_hydro.cpp")
    rs = flux + ("_mhd" if (mhd and flux in ("hlle", "llf", "roe")) else "")
    if mhd and flux == "hlld" and eos == "isothermal":
        rs += "_iso"                                    # configure.py:406-409
    rs += ".cpp"
    files += [os.path.join(src, "eos/general/noop.cpp"),
              os.path.join(src, "eos", eosf),
              os.path.join(src, "eos/eos_high_order.cpp"),
              os.path.join(src, "eos/eos_scalars.cpp"),
              os.path.join(src, "hydro/rsolvers", "mhd" if mhd else "hydro", rs),
              os.path.join(src, "pgen/default_pgen.cpp")]
    return files


def compile_one(args):
    srcf, objf, incs = args
    if os.path.exists(objf) and os.path.getmtime(objf) > os.path.getmtime(srcf):
        return objf, 0, ""
    cmd = ["g++"] + CXXFLAGS + incs + ["-c", srcf, "-o", objf]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return objf, r.returncode, r.stderr


def build(cfg, ref, jobs):
    mhd, flux, ng, pgens, nscalars, eos = cfg_tuple(cfg)
    src = os.path.join(ref, "src")
    root = os.path.join(OUT, cfg)
    inc = os.path.join(root, "inc")
    obj = os.path.join(root, "obj")
    os.makedirs(os.path.join(inc, "sub"), exist_ok=True)
    os.makedirs(obj, exist_ok=True)
    with open(os.path.join(src, "defs.hpp.in")) as f:
        text = f.read()
    for k, v in defs_for(mhd, flux, ng, nscalars, eos).items():
        text = text.replace("@" + k + "@", v)
    assert not re.search(r"@[A-Z0-9_]+@", text), "unfilled key in defs.hpp"
    dpath = os.path.join(inc, "defs.hpp")
    if not os.path.exists(dpath) or open(dpath).read() != text:
        with open(dpath, "w") as f:
            f.write(text)
    # "defs.hpp" resolves through -I inc ; "../defs.hpp" through -I inc/sub
    incs = ["-I", inc, "-I", os.path.join(inc, "sub")]
    files = source_list(src, mhd, flux, eos)
    work = []
    for s in files:
        rel = os.path.relpath(s, src).replace("/", "__")[:-4] + ".o"
        work.append((s, os.path.join(obj, rel), incs))
    # "local:<name>": a pgen of this repository (oracle/pgen/<name>.cpp, test infrastructure
    # written against the reference's ProblemGenerator / user-hook API), compiled with the
    # reference's pgen directory on the include path so that its "../athena.hpp" resolves
    def pgen_src(p):
        if p.startswith("local:"):
            return os.path.join(HERE, "pgen", p[6:] + ".cpp")
        return os.path.join(src, "pgen", p + ".cpp")
    local_inc = os.path.join(inc, "localpgen")
    os.makedirs(local_inc, exist_ok=True)
    pg_work = []
    for p in pgens:
        name = p.split(":")[-1]
        sfile = pgen_src(p)
        if p.startswith("local:"):
            # compile a one-line wrapper placed virtually inside src/pgen/
            wrap = os.path.join(local_inc, name + "_wrap.cpp")
            with open(wrap, "w") as f:
                f.write('#include "%s"\n' % sfile)
            # "../x.hpp" in the included file resolves relative to ITS directory first, then -I
            pg_work.append((wrap, os.path.join(obj, "pgen__" + name + ".o"),
                            incs + ["-I", os.path.join(src, "pgen")]))
        else:
            pg_work.append((sfile, os.path.join(obj, "pgen__" + name + ".o"), incs))
    failed = False
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        for objf, rc, err in ex.map(compile_one, work + pg_work):
            if rc != 0:
                failed = True
                sys.stderr.write("FAILED %s\n%s\n" % (objf, err))
    if failed:
        raise SystemExit("reference build failed for " + cfg)
    common = [w[1] for w in work]
    for p, w in zip(pgens, pg_work):
        exe = os.path.join(root, "athena_" + p.split(":")[-1])
        cmd = ["g++"] + CXXFLAGS + ["-s", "-o", exe] + common + [w[1]]   # -s: smaller to ship
        subprocess.run(cmd, check=True)
        print("built", exe)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("cfgs", nargs="*")
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(a.ref, "src")):
        print("reference tree not present at %s: keeping prebuilt oracle/_ref as is" % a.ref)
        return 0
    for cfg in (a.cfgs or list(CONFIGS)):
        build(cfg, a.ref, a.jobs)
    return 0


if __name__ == "__main__":
    sys.exit(main())
