"""CPU: the problem generators (athena-gamma_b200/pgen) against the initial state the
reference's own pgens produced (golden `init` dumps, active zones).  The pgens are host code
on the reference side too, so this only guards the synthetic bench inputs: tolerance 1e-12
relative (libm sin/cos of numpy vs glibc may differ in the last bit)."""
import types

import numpy as np
import pytest

import athena_gamma_b200 as ab
import oracle
import util
from athena_gamma_b200.mesh import MeshBlock


class FakeBlock:
    """MeshBlock stand-in whose coordinates come from the oracle (no GPU needed)"""
    shape = MeshBlock.shape

    def __init__(self, om, b, mhd):
        i = om.info[b]
        self.lid = self.gid = b
        self.lx1, self.lx2, self.lx3 = i["lx1"], i["lx2"], i["lx3"]
        self.ncells1, self.ncells2, self.ncells3 = i["nc1"], i["nc2"], i["nc3"]
        self.is_, self.ie, self.js, self.je, self.ks, self.ke = (i["is"], i["ie"], i["js"],
                                                                 i["je"], i["ks"], i["ke"])
        self.pmy_mesh = types.SimpleNamespace(mhd=mhd)
        self._om, self._b = om, b

    def coord(self, name):
        return np.array(self._om.array(self._b, name))


@pytest.mark.parametrize("name,pgen", [
    ("c5_blast_hlld_plm_vl2_8blk", "blast"),
    ("c2_linwave_hlld_plm_vl2_8blk", "linear_wave"),
    ("linwave_mhd_roe_plm_vl2_2blk", "linear_wave"),
    ("c3_ot_hlld_ppm_vl2_4blk", "orszag_tang"),
    ("c1_sod_hllc_plm_vl2_2blk", "shock_tube"),
])
def test_pgen_matches_reference_initial_state(name, pgen):
    g = util.Golden(name)
    p = oracle.params_from_athinput(g.par, g.mhd, g.solver, ng=g.ng)
    om = oracle.OracleMesh(p)
    pin = ab.ParameterInput()
    for blk, kv in g.par.items():
        for k, v in kv.items():
            pin.set(blk, k, v)
    for n, loc in enumerate(g.locs):
        b = om.block_of(*loc)
        pmb = FakeBlock(om, b, g.mhd)
        out = ab.pgen.BY_NAME[pgen](pmb, pin)
        K, J, I = slice(pmb.ks, pmb.ke + 1), slice(pmb.js, pmb.je + 1), slice(pmb.is_, pmb.ie + 1)
        ref = g.init[n]
        np.testing.assert_allclose(out["u"][:, K, J, I], ref["u"][:, K, J, I], rtol=1e-12,
                                   atol=1e-14, err_msg="%s u block %s" % (name, loc))
        if g.mhd:
            I1, J1, K1 = slice(pmb.is_, pmb.ie + 2), slice(pmb.js, pmb.je + 2), slice(pmb.ks, pmb.ke + 2)
            if pmb.ncells2 == 1:
                J1 = J
            if pmb.ncells3 == 1:
                K1 = K
            np.testing.assert_allclose(out["b1"][K, J, I1], ref["b1"][K, J, I1], rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(out["b2"][K, J1, I], ref["b2"][K, J1, I], rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(out["b3"][K1, J, I], ref["b3"][K1, J, I], rtol=1e-12, atol=1e-14)
