"""GPU: the CUDA path, called through the C ABI, against (a) the committed golden vectors of
the unmodified reference and (b) the C oracle stepping the same input.  Bar: bit-identical dt
sequence; conserved variables and face fields bit-identical (north_star allows 1e-12 relative;
we hold the stricter bit-exact bar and report the max deviation on failure)."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", util.golden_names())
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
