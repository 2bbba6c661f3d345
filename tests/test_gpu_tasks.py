"""GPU: every task-level entry point of the C ABI against the same task in the C oracle, on a
perturbed (non-smooth) state so that limiter branches, floors and all Riemann-solver branches
are exercised.  Bit-exact."""
import ctypes as C

import numpy as np
import pytest

import oracle
import util

pytestmark = pytest.mark.gpu

CASES = [
    # golden to borrow geometry from, xorder override, solver override
    ("c5_blast_hlld_plm_vl2_8blk", None, None),
    ("c5_blast_hlld_plm_vl2_1blk", 1, None),
    ("c3_ot_hlld_ppm_vl2_4blk", None, None),
    ("c4_kh_hllc_ppm_rk2_8blk", None, None),
    ("c1_sod_hllc_plm_vl2_2blk", None, None),
    ("linwave_mhd_roe_plm_vl2_2blk", None, None),
    ("linwave_mhd_hlle_plm_vl2", None, None),
    ("sod_roe_plm_vl2", None, None),
    ("sod_hlle_plm_vl2", None, None),
    ("blast_lhlld_plm_vl2_8blk", None, None),
    ("ot_lhlld_plm_vl2_4blk", None, None),
    ("blast_lhllc_plm_vl2_8blk", None, None),
    ("sod_lhllc_plm_vl2_2blk", None, None),
    ("blast_refl_hlld_plm_vl2_8blk", None, None),
    ("blast_mixedbc_hllc_plm_vl2_8blk", None, None),
    ("blast_hlld_ppm_rk3_8blk", None, None),
    # user-enrolled boundary function
    ("shkcloud2d_hllc_plm_vl2_4blk", None, None),
    ("shkcloud3d_hlld_plm_vl2_8blk", None, None),
    # 1-D / 2-D / non-cubic MHD
    ("bw1d_hlld_plm_vl2_2blk", None, None),
    ("bw2d_x2_hlld_plm_rk2_4blk", None, None),
    ("blast_noncubic_hlld_ppm_rk2_6blk", None, None),
    # characteristic reconstruction
    ("blast_hlld_plmc_vl2_8blk", None, None),
    ("blast_hllc_plmc_vl2_8blk", None, None),
    ("blast_hlld_ppmc_rk3_8blk", None, None),
    ("kh_hllc_ppmc_rk2_8blk", None, None),
    # LLF
    ("blast_llf_plm_vl2_8blk", None, None),
    ("blast_mhd_llf_plm_vl2_8blk", None, None),
    ("iso_blast_mhd_llf_plm_vl2_8blk", None, None),
    # user-enrolled explicit source function
    ("usersrc_lhllc_plm_vl2_8blk_s1", None, None),
    ("usersrc_hlld_plm_rk3_8blk", None, None),
    # constant-acceleration source term
    ("blast_grav_hllc_plm_vl2_8blk", None, None),
    ("blast_grav_hlld_plm_rk2_8blk", None, None),
    ("iso_blast_grav_hlle_plm_vl2_8blk", None, None),
    # isothermal EOS
    ("iso_khs_hlle_plm_vl2_4blk_s1", None, None),
    ("iso_blast_hlle_plm_vl2_8blk", None, None),
    ("iso_blast_hlle_plm_vl2_8blk", 1, None),
    ("iso_blast_mhd_hlld_plm_vl2_8blk", None, None),
    ("iso_ot_hlld_plm_rk2_4blk", None, None),
    ("iso_ot_mhd_hlle_plm_vl2_4blk", None, None),
    ("iso_blast_roe_plm_vl2_8blk", None, None),
    ("iso_blast_mhd_roe_plm_vl2_8blk", None, None),
    ("iso_kh2d_roe_plm_rk2_4blk", None, None),
    ("iso_ot_mhd_roe_plm_rk2_4blk", None, None),
    # passive scalars
    ("khs_lhllc_plm_vl2_4blk_s1", None, None),
    ("sods_lhllc_plm_vl2_2blk_s1", None, None),
    ("khs3d_hllc_ppm_rk3_8blk_s2", None, None),
    ("khs3d_mhd_hlld_plm_vl2_8blk_s1", None, None),
    ("khs3d_mhd_hlld_plm_vl2_8blk_s1", 1, None),
]


def perturbed(g, seed):
    """golden initial state with strong random perturbations (keeps rho, p > 0 mostly;
    a few cells get negative pressure to hit the floors)"""
    rng = np.random.default_rng(seed)
    init = []
    for blk in g.init:
        nb = {}
        u = blk["u"].copy()
        u[0] *= np.exp(rng.normal(0, 0.5, u[0].shape))
        u[1:4] += rng.normal(0, 0.5, u[1:4].shape) * u[0]
        if u.shape[0] > 4:          # adiabatic: energy; a few cells below the pressure floor
            u[4] *= np.exp(rng.normal(0, 0.3, u[4].shape))
            u[4] += 0.5 * (u[1] ** 2 + u[2] ** 2 + u[3] ** 2) / u[0]
            mask = rng.random(u[4].shape) < 0.002
            u[4][mask] = 1e-3
        else:                       # isothermal: a few cells below the density floor
            u[0][rng.random(u[0].shape) < 0.002] = 1e-30
        nb["u"] = u
        for f in g.fields[1:]:
            if f == "s":      # concentrations in [0, 1], a few below the floor
                c = rng.random(blk[f].shape)
                c[rng.random(c.shape) < 0.01] = -1e-3
                nb[f] = c * u[0]
            else:
                nb[f] = blk[f] + rng.normal(0, 0.3, blk[f].shape)
        init.append(nb)
    return init


@pytest.mark.parametrize("name,xorder,solver", CASES)
def test_task_by_task(name, xorder, solver):
    import gpu_util
    g = util.Golden(name)
    if xorder is not None:
        g.par["time"]["xorder"] = str(xorder)
    g.init = perturbed(g, 42)
    om = util.oracle_from_golden(g)          # exchange + cons2prim + bcs + dt on the oracle
    m = gpu_util.mesh_from_golden(g)
    m.initialize()
    L, h = m.L, m.h
    names_cc = ("u", "w") + (("bcc", "b1", "b2", "b3") if g.mhd else ())
    if g.nscalars:
        names_cc += ("s", "r")

    def compare(names, what):
        for pmb in m.my_blocks:
            b = om.block_of(pmb.lx1, pmb.lx2, pmb.lx3)
            for nm in names:
                util.assert_bitwise(pmb.get(nm), np.array(om.array(b, nm)),
                                    "%s: %s %s block %d" % (name, what, nm, pmb.gid))

    compare(names_cc, "initialize")
    assert m.dt == om.dt
    dt = om.dt
    xo = om.p.xorder
    for order in sorted({1, xo}):
        for pmb in m.my_blocks:
            b = om.block_of(pmb.lx1, pmb.lx2, pmb.lx3)
            om.L.ao_calc_fluxes(om.h, b, order)
            ab_check(L.ab_calc_fluxes(h, pmb.lid, order, dt), L)
            if g.mhd:
                om.L.ao_corner_e(om.h, b)
                ab_check(L.ab_corner_e(h, pmb.lid), L)
        compare_fluxes(m, om, g, name, order)
        if g.nscalars:
            for pmb in m.my_blocks:
                b = om.block_of(pmb.lx1, pmb.lx2, pmb.lx3)
                om.L.ao_calc_scalar_fluxes(om.h, b, order)
                ab_check(L.ab_calc_scalar_fluxes(h, pmb.lid, order), L)
            compare_scalar_fluxes(m, om, name, order)
        if g.mhd:
            compare_interior_emf(m, om, name, "corner_e order %d" % order)
            om.L.ao_emf_exchange(om.h)
            ab_check(L.ab_emf_exchange(h), L)
            compare_interior_emf(m, om, name, "emf exchange order %d" % order)
    # IntegrateHydro / IntegrateField, task by task (stage-1 VL2 weights)
    w = (C.c_double * 5)(1.0, 1.0, 0.0, 0.0, 0.0)
    for pmb in m.my_blocks:
        b = om.block_of(pmb.lx1, pmb.lx2, pmb.lx3)
        om.L.ao_zero_reg1(om.h, b)
        om.L.ao_weighted_ave_cc(om.h, b, 1, 0, w)
        om.L.ao_swap_cc(om.h, b)
        om.L.ao_add_flux_div(om.h, b, 0.5 * dt)
        ab_check(L.ab_zero(h, pmb.lid, 1), L)
        ab_check(L.ab_weighted_ave(h, pmb.lid, 1, 0, w), L)
        ab_check(L.ab_swap(h, pmb.lid, 0), L)
        ab_check(L.ab_add_flux_div(h, pmb.lid, 0.5 * dt), L)
        om.L.ao_add_source_terms(om.h, b, 0.25, 0.5 * dt)   # SRC_TERM (no-op without sources)
        ab_check(L.ab_add_source_terms(h, pmb.lid, 0.25, 0.5 * dt), L)
        if g.mhd:
            om.L.ao_weighted_ave_fc(om.h, b, 1, 0, w)
            om.L.ao_swap_fc(om.h, b)
            om.L.ao_ct(om.h, b, 0.5 * dt)
            ab_check(L.ab_zero(h, pmb.lid, 7), L)
            ab_check(L.ab_weighted_ave(h, pmb.lid, 7, 4, w), L)
            ab_check(L.ab_swap(h, pmb.lid, 4), L)
            ab_check(L.ab_ct(h, pmb.lid, 0.5 * dt), L)
    if g.nscalars:
        # IntegrateScalars: the oracle runs the whole task (stage 1), the device task by task
        beta0 = 0.5 if g.par["time"].get("integrator", "vl2") == "vl2" else 1.0
        S, S1 = 25, 26
        for pmb in m.my_blocks:
            b = om.block_of(pmb.lx1, pmb.lx2, pmb.lx3)
            om.L.ao_integrate_scalars(om.h, b, 1)
            ab_check(L.ab_zero(h, pmb.lid, S1), L)
            ab_check(L.ab_weighted_ave(h, pmb.lid, S1, S, w), L)
            ab_check(L.ab_swap(h, pmb.lid, S), L)
            ab_check(L.ab_add_scalar_flux_div(h, pmb.lid, beta0 * dt), L)
        om.L.ao_exchange_scalars(om.h)
    om.L.ao_exchange_cc(om.h)
    om.L.ao_exchange_fc(om.h)
    ab_check(L.ab_bvals_exchange(h), L)
    compare(("u", "u1") + (("b1", "b2", "b3", "b1_1", "b1_2", "b1_3") if g.mhd else ())
            + (("s", "s1") if g.nscalars else ()), "integrate+exchange")
    for pmb in m.my_blocks:
        b = om.block_of(pmb.lx1, pmb.lx2, pmb.lx3)
        om.L.ao_primitives(om.h, b)
        om.L.ao_physical_bcs(om.h, b)
        ab_check(L.ab_primitives(h, pmb.lid), L)
        ab_check(L.ab_physical_bcs(h, pmb.lid), L)
    compare(names_cc, "primitives+bcs")
    for pmb in m.my_blocks:
        b = om.block_of(pmb.lx1, pmb.lx2, pmb.lx3)
        d = C.c_double()
        ab_check(L.ab_new_block_dt(h, pmb.lid, C.byref(d)), L)
        assert d.value == om.L.ao_new_block_dt(om.h, b), "new_block_dt block %d" % pmb.gid


def ab_check(rc, L):
    assert rc == 0, L.ab_last_error().decode()


def face_region(pmb, g, d):
    """index ranges the reference's CalculateFluxes writes (calculate_fluxes.cpp:62-74 ...)"""
    is_, ie, js, je, ks, ke = pmb.is_, pmb.ie, pmb.js, pmb.je, pmb.ks, pmb.ke
    f2, f3 = pmb.ncells2 > 1, pmb.ncells3 > 1
    if d == 0:
        i, j, k = (is_, ie + 1), (js, je), (ks, ke)
        if g.mhd and f2:
            j = (js - 1, je + 1)
            if f3:
                k = (ks - 1, ke + 1)
    elif d == 1:
        i, j, k = (is_ - 1, ie + 1), (js, je + 1), (ks, ke)
        if g.mhd and f3:
            k = (ks - 1, ke + 1)
    else:
        i, j, k = (is_, ie), (js, je), (ks, ke + 1)
        if g.mhd:
            i, j = (is_ - 1, ie + 1), (js - 1, je + 1)
    return (slice(k[0], k[1] + 1), slice(j[0], j[1] + 1), slice(i[0], i[1] + 1))


def compare_fluxes(m, om, g, name, order):
    ndim = 1 + (m.params.nx2 > 1) + (m.params.nx3 > 1)
    for pmb in m.my_blocks:
        b = om.block_of(pmb.lx1, pmb.lx2, pmb.lx3)
        for d in range(ndim):
            sl = face_region(pmb, g, d)
            fo = np.array(om.array(b, "flux%d" % (d + 1)))
            fg = pmb.get("flux%d" % (d + 1))
            util.assert_bitwise(fg[(slice(None),) + sl], fo[(slice(None),) + sl],
                                "%s: flux%d order %d block %d" % (name, d + 1, order, pmb.gid))
            if g.mhd:
                names = [("e3_x1f", "e2_x1f", "wght1"), ("e1_x2f", "e3_x2f", "wght2"),
                         ("e2_x3f", "e1_x3f", "wght3")][d]
                for nm in names:
                    util.assert_bitwise(pmb.get(nm)[sl], np.array(om.array(b, nm))[sl],
                                        "%s: %s order %d block %d" % (name, nm, order, pmb.gid))


def compare_scalar_fluxes(m, om, name, order):
    """faces PassiveScalars::AddFluxDivergence reads (the only ones the product evaluates)"""
    ndim = 1 + (m.params.nx2 > 1) + (m.params.nx3 > 1)
    for pmb in m.my_blocks:
        b = om.block_of(pmb.lx1, pmb.lx2, pmb.lx3)
        for d in range(ndim):
            hi = [pmb.ke, pmb.je, pmb.ie]
            hi[2 - d] += 1
            sl = (slice(None), slice(pmb.ks, hi[0] + 1), slice(pmb.js, hi[1] + 1),
                  slice(pmb.is_, hi[2] + 1))
            nm = "sflux%d" % (d + 1)
            util.assert_bitwise(pmb.get(nm)[sl], np.array(om.array(b, nm))[sl],
                                "%s: %s order %d block %d" % (name, nm, order, pmb.gid))


def compare_interior_emf(m, om, name, what):
    for pmb in m.my_blocks:
        b = om.block_of(pmb.lx1, pmb.lx2, pmb.lx3)
        is_, ie, js, je, ks, ke = pmb.is_, pmb.ie, pmb.js, pmb.je, pmb.ks, pmb.ke
        f2, f3 = pmb.ncells2 > 1, pmb.ncells3 > 1
        K1 = slice(ks, ke + 2)
        J1 = slice(js, je + 2)
        rng = {"e1": (K1, J1, slice(is_, ie + 1)), "e2": (K1, slice(js, je + 1), slice(is_, ie + 2)),
               "e3": (slice(ks, ke + 1), J1, slice(is_, ie + 2))}
        for nm in ("e1", "e2", "e3"):
            if nm == "e1" and not f2:
                continue
            util.assert_bitwise(pmb.get(nm)[rng[nm]], np.array(om.array(b, nm))[rng[nm]],
                                "%s: %s %s block %d" % (name, what, nm, pmb.gid))
