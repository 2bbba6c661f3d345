"""CPU: the product's static-mesh-refinement EXECUTION path against the oracle, step by step.

What runs here is product code: the planner's rows (csrc/ab_smr.cpp), their interpretation
(csrc/ab_smr_exec.h -- the same templates ab_mesh.cu instantiates with CUDA launches) and the
per-cell bodies of the SMR kernels (csrc/ab_smr_cells.cuh -- the same functions the __global__
wrappers call), compiled for the host by tests/hostcheck/smr_host.cpp and pointed at the arrays
of an oracle mesh.  A second oracle mesh performs the same step with the oracle's own code
(which reproduces the reference's SMR runs bit for bit).  After every step -- ghost exchange
between levels, ProlongateBoundaries, flux correction -- all arrays of all MeshBlocks must be
identical.  Only the CUDA launch glue of the device path is left for the GPU test
(tests/test_gpu_smr.py)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
import util

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import athena_gamma_b200 as ab  # noqa: E402
from test_gpu_smr import device_smr_goldens  # noqa: E402
from test_smr_plan_cpu import _case_from_golden  # noqa: E402

DP = C.POINTER(C.c_double)
P2C = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)


@pytest.fixture(scope="module")
def smr_host():
    so = os.path.join(HERE, "hostcheck", "libsmr_host.so")
    src = os.path.join(HERE, "hostcheck", "smr_host.cpp")
    deps = [src] + [os.path.join(ROOT, "athena-gamma_b200", "csrc", f) for f in
                    ("ab_smr_cells.cuh", "ab_smr_exec.h", "ab_physics.cuh", "ab_types.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                        "-x", "c++", src, "-o", so], check=True)
    L = C.CDLL(so)
    L.hc_smr_run.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                             C.POINTER(C.c_int), C.POINTER(C.c_long), C.c_long, C.c_double,
                             C.c_double, C.c_double, C.c_double, C.c_int, P2C, C.c_void_p]
    return L


def product_rows(case):
    p = ab.lib.AbMeshParams()
    p.nx1, p.nx2, p.nx3 = case["nx"]
    p.bx1, p.bx2, p.bx3 = case["bx"]
    for k, v in case["lim"].items():
        setattr(p, k, v)
    for i, f in enumerate(case["bc"]):
        p.bc[i] = ab.lib.BC[f]
    p.nghost, p.nranks = case["ng"], 1
    regs = (ab.lib.AbRefinementRegion*len(case["regions"]))()
    for i, r in enumerate(case["regions"]):
        (regs[i].x1min, regs[i].x1max, regs[i].x2min, regs[i].x2max, regs[i].x3min,
         regs[i].x3max, regs[i].level) = r
    L = ab.lib.load()
    h = C.c_void_p()
    assert L.ab_smr_plan_create(C.byref(p), regs, len(regs), C.byref(h)) == 0
    n = L.ab_smr_plan_transfers(h, None, 0)
    rows = (C.c_long*(12*n))()
    L.ab_smr_plan_transfers(h, rows, n)
    L.ab_smr_plan_destroy(h)
    return rows, n


ARRAYS = ("u", "s", "w", "r", "flux1", "flux2", "flux3", "sflux1", "sflux2", "sflux3", "coarse_u",
          "coarse_w", "coarse_s", "coarse_r", "dx1f", "dx2f", "dx3f", "x1v", "x2v", "x3v", "cx1v",
          "cx2v", "cx3v")
STATE = ARRAYS[:14]


def _ptr(om, b, name):
    n = C.c_long()
    p = om.L.ao_array(om.h, b, name.encode(), C.byref(n))
    return C.cast(p, C.c_void_p).value if (p and n.value) else None


def _flat(om, b, name):
    n = C.c_long()
    p = om.L.ao_array(om.h, b, name.encode(), C.byref(n))
    return np.ctypeslib.as_array(p, shape=(n.value,)) if (p and n.value) else None


@pytest.mark.parametrize("name", device_smr_goldens())
def test_product_smr_steps_match_oracle(smr_host, name):
    g = util.Golden(name)
    case = _case_from_golden(name)
    rows, nrows = product_rows(case)
    A, B = util.oracle_from_golden(g), util.oracle_from_golden(g)   # identical, initialised
    OL = oracle.lib()
    OL.ao_smr_step.argtypes = [C.c_void_p, C.c_int]
    nb = A.nb
    i0 = A.info[0]
    ndim = 1 + (i0["nc2"] > 1) + (i0["nc3"] > 1)
    cnc = [len(_flat(A, 0, "cx%dv" % d)) for d in (1, 2, 3)]
    cng = (g.ng + 1)//2 + 1
    cs = [cng if n > 1 else 0 for n in cnc]
    dims = (C.c_int*22)(i0["nc1"], i0["nc2"], i0["nc3"], cnc[0], cnc[1], cnc[2], i0["is"],
                        i0["js"], i0["ks"], cs[0], cs[1], cs[2], ndim, 5, g.nscalars, g.ng,
                        case["bx"][0], case["bx"][1], case["bx"][2], i0["ie"], i0["je"], i0["ke"])
    root_level = min(i["level"] for i in A.info)
    nrb = [case["nx"][d]//case["bx"][d] for d in range(3)]
    bcs = (C.c_int*(6*nb))()
    for b, i in enumerate(A.info):
        dl = i["level"] - root_level
        for d, key in enumerate(("lx1", "lx2", "lx3")):
            flag_in, flag_out = oracle.BC[case["bc"][2*d]], oracle.BC[case["bc"][2*d+1]]
            physical = case["nx"][d] == 1
            bcs[6*b + 2*d] = flag_in if (physical or i[key] == 0) else -1
            bcs[6*b + 2*d + 1] = flag_out if (physical or i[key] == (nrb[d] << dl) - 1) else -1
    ptrs = (C.c_void_p*(23*nb))(*[_ptr(A, b, nm) for b in range(nb) for nm in ARRAYS])

    def p2c(user, lid, il, iu, jl, ju, kl, ku):
        A.L.ao_prim2cons(A.h, lid, il, iu, jl, ju, kl, ku)
        if g.nscalars:
            A.L.ao_scalar_prim2cons(A.h, lid, il, iu, jl, ju, kl, ku)
    cb = P2C(p2c)
    gamma = float(g.par["hydro"]["gamma"])

    def run_product(what):     # the floors the oracle mesh was built with
        smr_host.hc_smr_run(what, nb, ptrs, bcs, dims, rows, nrows, gamma, A.p.dfloor, A.p.pfloor,
                            A.p.sfloor, 0, cb, None)

    def same(step):
        for b in range(nb):
            for nm in STATE:
                a, o = _flat(A, b, nm), _flat(B, b, nm)
                if a is not None:
                    util.assert_bitwise(a, o, "%s after %s: block %d %s" % (name, step, b, nm))

    rng = np.random.default_rng(3)
    for rnd in range(2):
        # (round 0: the fixture's initial state; round 1: the same state perturbed, with garbage in
        # every ghost zone and coarse buffer, so that nothing can pass by having been right before)
        if rnd == 1:
            for b in range(nb):
                for nm in ("u", "s", "w", "r", "coarse_u", "coarse_w", "coarse_s", "coarse_r"):
                    a = _flat(A, b, nm)
                    if a is None:
                        continue
                    if nm in ("u", "s"):
                        a *= np.exp(rng.uniform(-0.2, 0.2, a.shape))
                    else:
                        a[...] = rng.uniform(0.5, 1.5, a.shape)
                    _flat(B, b, nm)[...] = a
        run_product(0); OL.ao_smr_step(B.h, 0); same("exchange")
        run_product(1); OL.ao_smr_step(B.h, 1); same("ProlongateBoundaries")
        for b in range(nb):
            for nm in ("flux1", "flux2", "flux3", "sflux1", "sflux2", "sflux3"):
                a = _flat(A, b, nm)
                if a is not None:
                    a[...] = rng.normal(0, 1, a.shape)
                    _flat(B, b, nm)[...] = a
        run_product(2); OL.ao_smr_step(B.h, 2); same("flux correction")
