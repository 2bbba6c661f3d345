"""CPU: the C oracle (oracle/) against the committed golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  Bar: bit-exact dt sequence and state
(conserved variables and face fields of every MeshBlock, ghost zones included)."""
import pytest

import util


@pytest.mark.parametrize("name", util.golden_names(include_smr=True))
def test_oracle_reproduces_reference_golden(name):
    g = util.Golden(name)
    m = util.oracle_from_golden(g)
    # first dt comes out of Mesh::Initialize (NewBlockTimeStep + NewTimeStep)
    assert m.dt == g.dts[0], ("dt0", m.dt, g.dts[0])

    def check_history(c):
        # HistoryOutput row of cycle c: time, dt, sums printed with %24.16e (17 digits = exact)
        if g.hst is not None:
            assert g.hst[c, 0] == m.time and g.hst[c, 1] == m.dt
            h = m.history()      # columns after ours: user-enrolled history outputs (pgen)
            util.assert_bitwise(h, g.hst[c, 2:2 + len(h)], "%s history row %d" % (name, c))
    check_history(0)
    for c in range(g.ncycles):
        used = m.cycle()
        assert used == g.dts[c], ("cycle %d used dt" % c, used, g.dts[c])
        assert m.dt == g.dts[c + 1], ("cycle %d new dt" % c, m.dt, g.dts[c + 1])
        check_history(c + 1)
    assert m.time == g.final_time
    for n, loc in enumerate(g.locs):
        b = m.block_of(*loc)
        for f in g.fields:
            util.assert_bitwise(m.array(b, f), g.final[n][f], "%s block %s %s" % (name, loc, f))
