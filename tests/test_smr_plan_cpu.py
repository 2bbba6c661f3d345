"""CPU: the product's static-mesh-refinement planner (csrc/ab_smr.cpp through ab_smr_plan_*)
against the C oracle, whose SMR path reproduces the reference's runs bit for bit
(tests/golden/smr_*.npz, tests/test_oracle_golden.py).

Compared exactly, on the refined meshes of the golden fixtures and a few more:
  * the Z-ordered MeshBlock list (level, logical location) -- also against the reference's own
    block list stored in the fixtures;
  * every block's neighbour list (offsets, type, gid, level, finer-leaf indices) and nblevel;
  * the transfer plan of one ghost exchange (same level / to finer / to coarser boxes), the
    ProlongateBoundaries work list and the flux-correction pairs, as sets of index rows: the
    product derives them sender by sender like the reference, the oracle receiver by receiver.
The load balance is checked against Mesh::CalculateLoadBalance restated in the test."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import oracle
import util
from test_physics_hostcheck import hc  # noqa: F401  (fixture)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athena_gamma_b200 as ab  # noqa: E402

SMR_GOLDENS = [n for n in util.golden_names(include_smr=True) if n.startswith("smr_")]

EXTRA = {
    # three levels, periodic, 3-D, non-cubic root grid
    "three_levels_3d": dict(nx=(24, 16, 16), bx=(4, 4, 4), ng=2, bc=["periodic"]*6, regions=[
        (-0.2, 0.1, -0.1, 0.1, -0.1, 0.1, 2), (0.3, 0.5, 0.3, 0.5, -0.5, -0.3, 1)]),
    # region touching non-periodic mesh boundaries in 2-D, NGHOST = 4
    "edges_2d_ng4": dict(nx=(32, 16, 1), bx=(8, 8, 1), ng=4,
                         bc=["outflow", "reflecting", "reflecting", "outflow", "periodic", "periodic"],
                         regions=[(-0.5, -0.4, -0.5, -0.3, -0.5, 0.5, 2)]),
    "one_d": dict(nx=(64, 1, 1), bx=(4, 1, 1), ng=2, bc=["outflow", "outflow"] + ["periodic"]*4,
                  regions=[(0.1, 0.2, -0.5, 0.5, -0.5, 0.5, 3)]),
}


def _case_from_golden(name):
    g = util.Golden(name)
    mesh, mb = g.par["mesh"], g.par.get("meshblock", {})
    nx = tuple(int(mesh.get("nx%d" % d, 1)) for d in (1, 2, 3))
    bx = tuple(int(mb.get("nx%d" % d, nx[d - 1])) for d in (1, 2, 3))
    lim = {k: float(mesh.get(k, dflt)) for k, dflt in (("x1min", -0.5), ("x1max", 0.5),
           ("x2min", -0.5), ("x2max", 0.5), ("x3min", -0.5), ("x3max", 0.5))}
    regions = []
    for bn, blk in g.par.items():
        if bn.startswith("refinement"):
            regions.append(tuple(float(blk.get(k, lim[k])) for k in
                                 ("x1min", "x1max", "x2min", "x2max", "x3min", "x3max"))
                           + (int(blk["level"]),))
    bc = [mesh.get(k, "periodic") for k in ("ix1_bc", "ox1_bc", "ix2_bc", "ox2_bc", "ix3_bc",
                                            "ox3_bc")]
    return dict(nx=nx, bx=bx, ng=g.ng, bc=bc, regions=regions, lim=lim, locs=g.locs)


def _both(case, nranks=1):
    lim = case.get("lim", dict(x1min=-0.5, x1max=0.5, x2min=-0.5, x2max=0.5, x3min=-0.5,
                               x3max=0.5))
    # product planner
    p = ab.lib.AbMeshParams()
    p.nx1, p.nx2, p.nx3 = case["nx"]
    p.bx1, p.bx2, p.bx3 = case["bx"]
    for k, v in lim.items():
        setattr(p, k, v)
    for i, f in enumerate(case["bc"]):
        p.bc[i] = ab.lib.BC[f]
    p.nghost = case["ng"]
    p.nranks = nranks
    regs = (ab.lib.AbRefinementRegion*len(case["regions"]))()
    for i, r in enumerate(case["regions"]):
        (regs[i].x1min, regs[i].x1max, regs[i].x2min, regs[i].x2max, regs[i].x3min,
         regs[i].x3max, regs[i].level) = r
    L = ab.lib.load()
    h = C.c_void_p()
    rc = L.ab_smr_plan_create(C.byref(p), regs, len(case["regions"]), C.byref(h))
    assert rc == 0, L.ab_smr_last_error()
    # oracle
    q = oracle.AoParams()
    q.nx1, q.nx2, q.nx3 = case["nx"]
    q.bx1, q.bx2, q.bx3 = case["bx"]
    for k, v in lim.items():
        setattr(q, k, v)
    for i, f in enumerate(case["bc"]):
        q.bc[i] = oracle.BC[f]
    q.ng = case["ng"]
    q.solver, q.xorder, q.gamma, q.cfl, q.tlim = oracle.SOLVER["hllc"], 2, 1.4, 0.3, 1.0
    q.dfloor = q.pfloor = q.sfloor = oracle.DEFAULT_FLOOR
    q.nref = len(case["regions"])
    for i, r in enumerate(case["regions"]):
        for c in range(6):
            q.ref[i][c] = r[c]
        q.ref_level[i] = r[6]
    om = oracle.OracleMesh(q)
    return L, h, om


def _plan_blocks(L, h):
    n = L.ab_smr_plan_nblocks(h)
    rows = (C.c_long*(5*n))()
    assert L.ab_smr_plan_blocks(h, rows, n) == n
    return np.array(rows).reshape(n, 5)


ALL = [("golden:" + n, None) for n in SMR_GOLDENS] + [("extra:" + k, v) for k, v in EXTRA.items()]


@pytest.mark.parametrize("name,case", ALL, ids=[a for a, _ in ALL])
def test_block_list_neighbours_and_plan_match_oracle(name, case):
    if case is None:
        case = _case_from_golden(name.split(":", 1)[1])
    L, h, om = _both(case)
    try:
        blocks = _plan_blocks(L, h)
        mine = [(int(r[1]), int(r[2]), int(r[3]), int(r[0])) for r in blocks]
        theirs = [(i["lx1"], i["lx2"], i["lx3"], i["level"]) for i in om.info]
        assert mine == theirs
        if "locs" in case:      # the reference's own block list (restart file of the fixture)
            assert mine == [tuple(l) for l in case["locs"]]
        assert len(set(r[0] for r in blocks)) >= 2, "mesh is not refined"
        OL = oracle.lib()
        OL.ao_neighbors.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        for g in range(len(blocks)):
            a, b = (C.c_int*(8*56))(), (C.c_int*(8*56))()
            la, lb = (C.c_int*27)(), (C.c_int*27)()
            na = L.ab_smr_plan_neighbors(h, g, a, la)
            nb = OL.ao_neighbors(om.h, g, b, lb)
            assert na == nb, (g, na, nb)
            assert list(a)[:8*na] == list(b)[:8*nb], "neighbour list of block %d" % g
            assert list(la) == list(lb), "nblevel of block %d" % g
        # transfer plan
        n = L.ab_smr_plan_transfers(h, None, 0)
        rows = (C.c_long*(12*n))()
        L.ab_smr_plan_transfers(h, rows, n)
        mine = sorted(tuple(r) for r in np.array(rows).reshape(n, 12).tolist())
        OL.ao_smr_transfers.restype = C.c_long
        OL.ao_smr_transfers.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.c_long]
        m = OL.ao_smr_transfers(om.h, None, 0)
        orows = (C.c_long*(12*m))()
        OL.ao_smr_transfers(om.h, orows, m)
        theirs = sorted(tuple(r) for r in np.array(orows).reshape(m, 12).tolist())
        assert n == m and n > 0
        kinds = set(r[0] for r in mine)
        assert {0, 1, 2, 10, 11, 20} <= kinds
        assert mine == theirs
    finally:
        L.ab_smr_plan_destroy(h)


def test_load_balance_of_a_refined_mesh():
    from test_multirank_cpu import reference_load_balance
    case = EXTRA["three_levels_3d"]
    for nranks in (1, 2, 3, 8):
        L, h, _ = _both(case, nranks=nranks)
        try:
            blocks = _plan_blocks(L, h)
            assert list(blocks[:, 4]) == reference_load_balance(len(blocks), nranks)
        finally:
            L.ab_smr_plan_destroy(h)


def test_bad_refinement_is_rejected():
    L = ab.lib.load()
    p = ab.lib.AbMeshParams()
    p.nx1, p.nx2, p.nx3, p.bx1, p.bx2, p.bx3 = 16, 16, 1, 4, 4, 1
    p.x1min, p.x1max, p.x2min, p.x2max, p.x3min, p.x3max = -0.5, 0.5, -0.5, 0.5, -0.5, 0.5
    p.nghost = 3
    h = C.c_void_p()
    reg = (ab.lib.AbRefinementRegion*1)()
    reg[0].x1min, reg[0].x1max, reg[0].x2min, reg[0].x2max, reg[0].level = -0.1, 0.1, -0.1, 0.1, 1
    assert L.ab_smr_plan_create(C.byref(p), reg, 1, C.byref(h)) == ab.lib.AB_ERR_ARG
    assert b"even number of ghost" in L.ab_smr_last_error()
    p.nghost = 2
    reg[0].x1max = 0.7
    assert L.ab_smr_plan_create(C.byref(p), reg, 1, C.byref(h)) == ab.lib.AB_ERR_ARG
    assert b"smaller than the whole mesh" in L.ab_smr_last_error()


@pytest.mark.parametrize("gname", ["smr_blast3d_hllc_plm_vl2", "smr_blast2d_lvl2_bcs_hllc_plm_rk2",
                                   "smr_sod1d_hllc_plm_vl2"])
def test_device_restriction_and_prolongation_arithmetic(hc, gname):  # noqa: F811
    """ab_physics.cuh restrict_cc / prolong_grad / prolong_cc (the point functions the SMR
    kernels are built from), compiled for the host, against the oracle on a refined block."""
    g = util.Golden(gname)
    om = util.oracle_from_golden(g)
    OL = oracle.lib()
    IP, DP = C.POINTER(C.c_int), C.POINTER(C.c_double)
    OL.ao_smr_restrict_box.argtypes = [C.c_void_p, C.c_int, IP]
    OL.ao_smr_prolong_box.argtypes = [C.c_void_p, C.c_int, IP]
    hc.hc_smr_restrict.argtypes = [IP, DP, DP, DP, DP, DP, IP]
    hc.hc_smr_prolong.argtypes = [IP, DP, DP, DP, DP, DP, DP, DP, DP, IP]
    rng = np.random.default_rng(5)
    b = max(range(om.nb), key=lambda n: om.info[n]["level"])     # a refined block
    i = om.info[b]
    ndim = 1 + (i["nc2"] > 1) + (i["nc3"] > 1)
    cu = om.array(b, "coarse_u")
    nh = 5
    cnc = [len(om.array(b, "cx%dv" % d)) for d in (1, 2, 3)]
    cng = (g.ng + 1)//2 + 1
    cs = [cng if n > 1 else 0 for n in cnc]
    dims = (C.c_int*13)(i["nc1"], i["nc2"], i["nc3"], cnc[0], cnc[1], cnc[2], i["is"], i["js"],
                        i["ks"], cs[0], cs[1], cs[2], ndim)
    dp = lambda a: np.ascontiguousarray(a).ctypes.data_as(DP)   # noqa: E731
    # restriction over all coarse cells whose fine cells exist (active + ghosts)
    u = om.array(b, "u")
    u[...] = np.exp(rng.uniform(-1, 1, u.shape))
    box = (C.c_int*6)(*sum(([c - (g.ng//2 if n > 1 else 0), (n - 1 - c) + (g.ng//2 if n > 1 else 0)]
                            for c, n in zip(cs, cnc)), []))
    OL.ao_smr_restrict_box(om.h, b, box)
    want = np.array(om.array(b, "coarse_u")).reshape(nh, cnc[2], cnc[1], cnc[0])
    dx = [np.array(om.array(b, "dx%df" % d)) for d in (1, 2, 3)]
    for v in range(nh):
        got = np.zeros((cnc[2], cnc[1], cnc[0]))
        hc.hc_smr_restrict(dims, dp(dx[0]), dp(dx[1]), dp(dx[2]), dp(u[v]), dp(got), box)
        sl = tuple(slice(box[2*d], box[2*d+1] + 1) for d in (2, 1, 0))
        util.assert_bitwise(got[sl], want[v][sl], "restriction var %d" % v)
    # prolongation of a coarse box that has a one-cell margin
    cw = om.array(b, "coarse_w")
    cw[...] = rng.normal(0, 1, cw.shape)
    pbox = (C.c_int*6)(*sum(([c - (1 if n > 1 else 0), (n - 1 - c) + (1 if n > 1 else 0)]
                             for c, n in zip(cs, cnc)), []))
    w = om.array(b, "w")
    w[...] = 0.0
    OL.ao_smr_prolong_box(om.h, b, pbox)
    wantw = np.array(om.array(b, "w"))
    xv = [np.array(om.array(b, "x%dv" % d)) for d in (1, 2, 3)]
    cxv = [np.array(om.array(b, "cx%dv" % d)) for d in (1, 2, 3)]
    cw4 = np.array(cw).reshape(nh, cnc[2], cnc[1], cnc[0])
    for v in range(nh):
        got = np.zeros(w.shape[1:])
        hc.hc_smr_prolong(dims, dp(xv[0]), dp(xv[1]), dp(xv[2]), dp(cxv[0]), dp(cxv[1]),
                          dp(cxv[2]), dp(cw4[v]), dp(got), pbox)
        util.assert_bitwise(got, wantw[v], "prolongation var %d" % v)


@pytest.mark.parametrize("gname", ["smr_blast3d_refl_lhllc_plm_rk3", "smr_blast2d_lvl2_bcs_hllc_plm_rk2",
                                   "smr_sod1d_hllc_plm_vl2", "smr_khs3d_hllc_ppm_rk3_ng4_s2"])
def test_refined_mesh_host_setup_matches_oracle(gname):
    """ab_plan_create_refined = the host half of ab_mesh_create_refined: per-level block extents,
    coordinates (also in the mirrored ghost zones of reflecting boundaries), the coarse buffers'
    cell centres, levels and block order, against the oracle's arrays."""
    g = util.Golden(gname)
    case = _case_from_golden(gname)
    om = util.oracle_from_golden(g)
    p = ab.lib.AbMeshParams()
    p.nx1, p.nx2, p.nx3 = case["nx"]
    p.bx1, p.bx2, p.bx3 = case["bx"]
    for k, v in case["lim"].items():
        setattr(p, k, v)
    for i, f in enumerate(case["bc"]):
        p.bc[i] = ab.lib.BC[f]
    p.nghost, p.nranks, p.xorder = case["ng"], 1, 2
    p.solver, p.gamma, p.cfl_number, p.tlim = ab.lib.SOLVER["hllc"], 1.4, 0.3, 1.0
    regs = (ab.lib.AbRefinementRegion*len(case["regions"]))()
    for i, r in enumerate(case["regions"]):
        (regs[i].x1min, regs[i].x1max, regs[i].x2min, regs[i].x2max, regs[i].x3min,
         regs[i].x3max, regs[i].level) = r
    L = ab.lib.load()
    h = C.c_void_p()
    rc = L.ab_plan_create_refined(C.byref(p), regs, len(regs), C.byref(h))
    assert rc == 0, L.ab_last_error()
    try:
        nb = L.ab_mesh_nblocks_local(h)
        assert nb == om.nb == L.ab_mesh_nblocks_total(h)
        DP = C.POINTER(C.c_double)
        for lid in range(nb):
            info = (C.c_long*13)()
            assert L.ab_block_info(h, lid, info) == 0
            i = om.info[lid]
            assert (info[1], info[2], info[3]) == (i["lx1"], i["lx2"], i["lx3"])
            assert L.ab_block_level(h, lid) == i["level"]
            for d in range(3):
                for what, nm in ((0, "x%df"), (1, "x%dv"), (2, "dx%df"), (8, "cx%dv")):
                    want = np.array(om.array(lid, nm % (d + 1)))
                    n = L.ab_plan_geometry(h, lid, d, what, None, 0)
                    got = np.zeros(n)
                    L.ab_plan_geometry(h, lid, d, what, got.ctypes.data_as(DP), n)
                    util.assert_bitwise(got, want, "block %d %s" % (lid, nm % (d + 1)))
    finally:
        L.ab_mesh_destroy(h)
