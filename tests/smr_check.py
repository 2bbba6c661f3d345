"""Run by tests/test_gpu_smr.py in its own process: the device path on statically refined
meshes (ab_mesh_create_refined) against the reference's golden vectors -- dt sequence and the
conserved variables (and scalars) of every MeshBlock on every level, ghost zones included,
bit for bit."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    sys.path.insert(0, p)
import gpu_util  # noqa: E402
import util  # noqa: E402


def main(names):
    failed = []
    for name in names:
        g = util.Golden(name)
        try:
            m = gpu_util.mesh_from_golden(g)
            assert m.nbtotal == len(g.locs), (m.nbtotal, len(g.locs))
            m.initialize()
            assert m.dt == g.dts[0], ("dt0", m.dt, g.dts[0])
            dts = m.cycles(g.ncycles)
            assert list(dts) == list(g.dts[:g.ncycles]), (list(dts), list(g.dts))
            assert m.dt == g.dts[g.ncycles] and m.time == g.final_time
            for n, loc in enumerate(g.locs):
                pmb = m.block_of(*loc)
                for f in g.fields:
                    util.assert_bitwise(pmb.get(f), g.final[n][f], "%s block %s %s" % (name, loc, f))
            print("ok", name, flush=True)
        except Exception as ex:      # keep going: report every fixture
            failed.append(name)
            print("FAILED", name, repr(ex)[:400], flush=True)
    print("smr done: %d failed" % len(failed))
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
