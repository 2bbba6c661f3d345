"""CPU: nonuniform (geometric) mesh spacing, mesh/x?rat != 1.

 * the product's host-side geometry (ab_plan_geometry on a host-only plan: mesh generator,
   block extents, dx?f, the cell-centred-field weights) against the oracle, which is pinned to
   the unmodified reference by the *_nonuni_* goldens (tests/test_oracle_golden.py);
 * the product's device reconstruction functions (ab_physics.cuh: plm_nu / ppm_nu and the
   characteristic variants, compiled for the host by tests/hostcheck) fed with the product's
   own geometry table, against the oracle's nonuniform PLM / PPM on the same lines.
Bit-exact.  The GPU parity tests (tests/test_gpu_golden.py, *_nonuni_* fixtures) then cover the
kernels."""
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

DP = C.POINTER(C.c_double)


def _dp(a):
    return a.ctypes.data_as(DP)


def plan_geometry(plan, lid, d, what):
    n = plan.L.ab_plan_geometry(plan.h, lid, d, what, None, 0)
    assert n >= 0
    out = np.zeros(n)
    if n:
        plan.L.ab_plan_geometry(plan.h, lid, d, what, _dp(out), n)
    return out


CASES = [
    # athinput, overrides, mhd, flux, nghost
    ("athinput.blast", ["mesh/nx1=16", "mesh/nx2=16", "mesh/nx3=16", "meshblock/nx1=8",
                        "meshblock/nx2=8", "meshblock/nx3=8", "mesh/x1rat=1.05",
                        "mesh/x2rat=0.96", "mesh/x3rat=1.03"], True, "hlld", 2),
    ("athinput.blast", ["mesh/nx1=24", "mesh/nx2=12", "mesh/nx3=8", "meshblock/nx1=12",
                        "meshblock/nx2=6", "meshblock/nx3=4", "mesh/x1rat=0.95",
                        "mesh/x3rat=1.06", "time/xorder=3", "mesh/ix1_bc=reflecting",
                        "mesh/ox1_bc=outflow", "mesh/ix3_bc=outflow",
                        "mesh/ox3_bc=reflecting"], True, "hlld", 3),
    ("athinput.orszag_tang", ["mesh/nx1=32", "mesh/nx2=32", "meshblock/nx1=16",
                              "meshblock/nx2=16", "mesh/x2rat=1.07"], True, "hlld", 3),
    ("athinput.sod", ["mesh/nx1=64", "meshblock/nx1=16", "mesh/x1rat=1.03"], False, "hllc", 2),
]


def _both(case):
    inp, ov, mhd, flux, ng = case
    pin = ab.ParameterInput(path=os.path.join(ROOT, "inputs", inp))
    for o in ov:      # (ModifyFromCmdline only changes existing keys; x?rat are new ones)
        bk, v = o.split("=", 1)
        pin.set(*bk.split("/", 1), v)
    plan = ab.MeshPlan(pin, mhd, flux, nghost=ng)
    par = oracle_par(pin)
    om = oracle.OracleMesh(oracle.params_from_athinput(par, mhd, flux, ng=ng))
    return plan, om


def oracle_par(pin):
    return {b: dict(kv) for b, kv in pin.blocks.items()}


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0] + ":" + ",".join(
    o for o in c[1] if "rat" in o))
def test_geometry_matches_oracle(case):
    plan, om = _both(case)
    assert plan.nblocal == om.nb
    for lid, blk in enumerate(plan.my_blocks):
        b = om.block_of(blk.lx1, blk.lx2, blk.lx3)
        for d in range(3):
            for what, name in ((0, "x%df"), (1, "x%dv"), (2, "dx%df")):
                mine = plan_geometry(plan, lid, d, what)
                util.assert_bitwise(mine, om.array(b, name % (d + 1)),
                                    "block %d %s" % (lid, name % (d + 1)))
            # CalculateCellCenteredField weights
            xf, xv, dxf = (om.array(b, n % (d + 1)) for n in ("x%df", "x%dv", "dx%df"))
            nc = len(xv)
            if nc == 1:
                continue
            info = om.info[b]
            s, e = info[("is", "js", "ks")[d]], info[("ie", "je", "ke")[d]]
            nonuni = int(om.p.xrat[d] != 1.0)
            lw, rw = np.zeros(nc), np.zeros(nc)
            oracle.lib().ao_bcc_weights(d, nonuni, nc, s, e, om.p.ng, _dp(xf), _dp(xv), _dp(dxf),
                                        _dp(lw), _dp(rw))
            util.assert_bitwise(plan_geometry(plan, lid, d, 6), lw, "lw dir %d" % d)
            util.assert_bitwise(plan_geometry(plan, lid, d, 7), rw, "rw dir %d" % d)
            tab = plan_geometry(plan, lid, d, 5)
            assert len(tab) == (13*nc if nonuni else 0)


def _line_inputs(om, b, d, rng, nvar):
    xf, xv, dxf = (np.ascontiguousarray(om.array(b, n % (d + 1)))
                   for n in ("x%df", "x%dv", "dx%df"))
    nc = len(xv)
    q = np.exp(rng.uniform(-1, 1, (nvar, nc)))
    q[:, nc//2:] *= 3.0                      # a jump
    q[1] = rng.normal(0, 1, nc)              # sign changes / extrema
    if nc > 8:
        q[2, 3:7] = q[2, 3]                  # flat stretch
    return xf, xv, dxf, nc, np.ascontiguousarray(q)


@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("case", CASES[:3], ids=["blast3d", "blast_noncubic_ng3", "ot2d"])
def test_recon_line_matches_oracle(hc, case, order):  # noqa: F811
    hc.hc_recon_line.argtypes = [C.c_int]*3 + [DP, DP, DP, C.c_int, DP, C.c_int, C.c_int, DP, DP]
    plan, om = _both(case)
    if order == 3 and om.p.ng < 3:
        pytest.skip("PPM needs 3 ghost cells")
    rng = np.random.default_rng(7)
    checked = 0
    for lid, blk in enumerate(plan.my_blocks):
        b = om.block_of(blk.lx1, blk.lx2, blk.lx3)
        for d in range(3):
            nonuni = int(om.p.xrat[d] != 1.0)
            xf, xv, dxf, nc, q = _line_inputs(om, b, d, rng, 4)
            if nc == 1:
                continue
            info = om.info[b]
            s, e = info[("is", "js", "ks")[d]], info[("ie", "je", "ke")[d]]
            po, mo = np.zeros_like(q), np.zeros_like(q)
            oracle.lib().ao_recon_line(d, nonuni, order, nc, s, e, om.p.ng, _dp(xf), _dp(xv),
                                       _dp(dxf), 4, _dp(q), s - 1, e + 1, _dp(po), _dp(mo))
            wp, wm = plan_geometry(plan, lid, d, 3), plan_geometry(plan, lid, d, 4)
            tab = plan_geometry(plan, lid, d, 5)
            ph, mh = np.zeros_like(q), np.zeros_like(q)
            hc.hc_recon_line(d + 1 if nonuni else 0, order, nc, _dp(wp), _dp(wm),
                             _dp(tab) if nonuni else None, 4, _dp(q), s - 1, e + 1, _dp(ph),
                             _dp(mh))
            util.assert_bitwise(ph, po, "plus dir %d order %d" % (d, order))
            util.assert_bitwise(mh, mo, "minus dir %d order %d" % (d, order))
            checked += nonuni
    assert checked > 0


@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("mhd", [False, True])
def test_recon_line_char_matches_oracle(hc, mhd, order):  # noqa: F811
    hc.hc_recon_line_char.argtypes = [C.c_int]*4 + [DP, DP, DP, DP, DP] + [C.c_double]*3 + \
        [C.c_int, C.c_int, DP, DP]
    plan, om = _both(CASES[1])
    rng = np.random.default_rng(11)
    gamma, fl = 5.0/3.0, oracle.DEFAULT_FLOOR
    for lid, blk in enumerate(plan.my_blocks[:2]):
        b = om.block_of(blk.lx1, blk.lx2, blk.lx3)
        for d in (0, 2):                      # x1rat and x3rat are set in this case
            xf, xv, dxf, nc, _ = _line_inputs(om, b, d, rng, 4)
            q = np.zeros((7, nc))
            q[0] = np.exp(rng.uniform(-1, 1, nc))
            q[1:4] = rng.normal(0, 1, (3, nc))
            q[4] = np.exp(rng.uniform(-1, 1, nc))
            q[5:7] = rng.normal(0, 1, (2, nc))
            bx = rng.normal(0, 1, nc)
            info = om.info[b]
            s, e = info[("is", "js", "ks")[d]], info[("ie", "je", "ke")[d]]
            po, mo = np.zeros_like(q), np.zeros_like(q)
            oracle.lib().ao_recon_line_char(d, 1, order, int(mhd), nc, s, e, om.p.ng, _dp(xf),
                                            _dp(xv), _dp(dxf), _dp(q), _dp(bx), gamma, fl, fl,
                                            s - 1, e + 1, _dp(po), _dp(mo))
            wp, wm = plan_geometry(plan, lid, d, 3), plan_geometry(plan, lid, d, 4)
            tab = plan_geometry(plan, lid, d, 5)
            ph, mh = np.zeros_like(q), np.zeros_like(q)
            hc.hc_recon_line_char(d + 1, order, int(mhd), nc, _dp(wp), _dp(wm), _dp(tab), _dp(q),
                                  _dp(bx), gamma, fl, fl, s - 1, e + 1, _dp(ph), _dp(mh))
            util.assert_bitwise(ph, po, "char plus dir %d order %d" % (d, order))
            util.assert_bitwise(mh, mo, "char minus dir %d order %d" % (d, order))
