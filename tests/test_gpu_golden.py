"""GPU: the CUDA path, called through the C ABI, against (a) the committed golden vectors of
the unmodified reference and (b) the C oracle stepping the same input.  Bar: bit-identical dt
sequence; conserved variables and face fields bit-identical (north_star allows 1e-12 relative;
we hold the stricter bit-exact bar and report the max deviation on failure)."""
import os

import numpy as np
import pytest

import util

HERE = os.path.dirname(os.path.abspath(__file__))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", util.gpu_params(util.golden_names()))
def test_cuda_reproduces_reference_golden(name):
    import gpu_util
    g = util.Golden(name)
    m = gpu_util.mesh_from_golden(g)
    m.initialize()
    assert m.dt == g.dts[0], ("dt0", m.dt, g.dts[0])
    dts = m.cycles(g.ncycles)
    assert len(dts) == g.ncycles
    for c in range(g.ncycles):
        assert dts[c] == g.dts[c], ("cycle %d dt" % c, dts[c], g.dts[c])
    assert m.dt == g.dts[g.ncycles]
    assert m.time == g.final_time
    for n, loc in enumerate(g.locs):
        pmb = m.block_of(*loc)
        for f in g.fields:
            util.assert_bitwise(pmb.get(f), g.final[n][f], "%s block %s %s" % (name, loc, f))


@pytest.mark.parametrize("name", ["c2_linwave_hlld_plm_vl2_8blk", "c4_kh_hllc_ppm_rk2_8blk",
                                  "c1_sod_hllc_plm_vl2_2blk"])
def test_cuda_matches_oracle_async_cycles(name):
    """Same, in the no-host-sync mode used by bench.py (dt never leaves the device)."""
    import gpu_util
    g = util.Golden(name)
    m = gpu_util.mesh_from_golden(g)
    m.initialize()
    om = util.oracle_from_golden(g)
    dts = m.cycles(g.ncycles, async_=True)
    odts = [om.cycle() for _ in range(g.ncycles)]
    assert list(dts) == odts
    for pmb in m.my_blocks:
        b = om.block_of(pmb.lx1, pmb.lx2, pmb.lx3)
        for f in g.fields + ("w",) + (("bcc",) if g.mhd else ()):
            util.assert_bitwise(pmb.get(f), np.array(om.array(b, f)), "%s %s" % (name, f))


@pytest.mark.parametrize("name", ["c5_blast_hlld_plm_vl2_8blk", "c3_ot_hlld_ppm_vl2_4blk",
                                  "blast_refl_hlld_plm_vl2_8blk", "khs3d_mhd_hlld_plm_vl2_8blk_s1",
                                  "c1_sod_hllc_plm_vl2_2blk", "iso_ot_hlld_plm_rk2_4blk",
                                  "shkcloud2d_hllc_plm_vl2_4blk"])
def test_overlapped_schedule_is_bit_identical(name, monkeypatch):
    """The multi-GPU schedule (EMF / ghost transfers on a second stream, ConservedToPrimitive
    split into active cells and ghost shell) forced on one GPU with AB_OVERLAP=1."""
    import gpu_util
    monkeypatch.setenv("AB_OVERLAP", "1")
    g = util.Golden(name)
    m = gpu_util.mesh_from_golden(g)
    m.initialize()
    dts = m.cycles(g.ncycles, async_=(not util.user_bcs_for(g)))
    assert list(dts) == list(g.dts[:g.ncycles])
    for n, loc in enumerate(g.locs):
        pmb = m.block_of(*loc)
        for f in g.fields:
            util.assert_bitwise(pmb.get(f), g.final[n][f], "%s block %s %s" % (name, loc, f))


def test_block_by_block_launches_are_bit_identical():
    """AB_NO_BATCH=1: one launch per task AND MeshBlock (the schedule before ab_batch.cuh, still
    what the task-level entry points and refined meshes use) against the same goldens the default
    one-launch-over-all-blocks cycle reproduces in test_cuda_reproduces_reference_golden."""
    import subprocess
    import sys
    names = ["c5_blast_hlld_plm_vl2_8blk", "c3_ot_hlld_ppm_vl2_4blk", "c4_kh_hllc_ppm_rk2_8blk",
             "blast_mixedbc_hllc_plm_vl2_8blk", "khs3d_mhd_hlld_plm_vl2_8blk_s1"]
    r = subprocess.run([sys.executable, os.path.join(HERE, "smr_check.py")] + names,
                       env=dict(os.environ, AB_NO_BATCH="1"), capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "smr done: 0 failed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name", ["usersrc_lhllc_plm_vl2_8blk_s1", "usersrc_hlld_plm_rk3_8blk",
                                  "usersrc_iso_hlle_plm_rk2_8blk"])
def test_device_side_user_source_function(name):
    """ab_enroll_user_explicit_source_function_device: the source term stays on the GPU (here as
    torch kernels on the library's stream over the library's registers, zero copies)."""
    import gpu_util
    import athena_gamma_b200 as ab
    g = util.Golden(name)
    pin = gpu_util.pin_from_par(g.par)
    m = ab.Mesh(pin, mhd=g.mhd, flux=g.solver, nghost=g.ng, nscalars=g.nscalars, eos=g.eos)
    m.enroll_user_explicit_source_function(gpu_util.central_gravity_source_torch(g.par),
                                           device=True)
    for n, loc in enumerate(g.locs):
        pmb = m.block_of(*loc)
        for f in g.fields:
            pmb.set(f, g.init[n][f])
    m.initialize()
    dts = m.cycles(g.ncycles)
    assert list(dts) == list(g.dts[:g.ncycles])
    for n, loc in enumerate(g.locs):
        pmb = m.block_of(*loc)
        for f in g.fields:
            util.assert_bitwise(pmb.get(f), g.final[n][f], "%s block %s %s" % (name, loc, f))


def history_close(h, ref, scale):
    """device tree sum vs the reference's running sum: |d| <= 1e-13 x sum of magnitudes"""
    assert len(h) == len(ref)
    tol = 1e-13*np.maximum(scale, 1e-300)
    bad = np.abs(h - ref) > tol
    assert not bad.any(), ("history", h, ref, tol)


@pytest.mark.parametrize("name", util.gpu_params([n for n in util.golden_names()
                                                  if util.Golden(n).hst is not None]))
def test_history_sums_match_reference_hst(name):
    """ab_history (on-device HistoryOutput sums) against the reference's .hst rows (17 digits)
    after every cycle.  The device sums in a fixed tree order, the reference keeps a running
    sum, so the bar is a tolerance: 1e-13 relative to the sum of |terms| (here bounded by the
    largest history value, all terms of one quantity having one sign or cancelling)."""
    import gpu_util
    g = util.Golden(name)
    m = gpu_util.mesh_from_golden(g)
    m.initialize()
    om = util.oracle_from_golden(g)
    for c in range(g.ncycles + 1):
        h = m.history()
        ref = g.hst[c, 2:2 + len(h)]
        assert g.hst[c, 0] == m.time and g.hst[c, 1] == m.dt
        util.assert_bitwise(om.history(), ref, "oracle history")
        history_close(h, ref, np.full(len(h), np.abs(ref).max()))
        if c < g.ncycles:
            m.cycles(1)
            om.cycle()
    # reproducible: the same state gives the same bits
    assert np.array_equal(m.history(), m.history())
