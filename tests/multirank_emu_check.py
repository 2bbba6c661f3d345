"""One rank of the CPU multi-rank check (tests/test_multirank_cpu.py starts WORLD_SIZE of these):
the emulated device path (AB_LIB = tests/hostcheck/libathena_b200_emu.so) with MeshBlocks sharded
over the ranks and the test-only NCCL stand-in (AB_NCCL_LIB = tests/hostcheck/libnccl_emu.so:
sockets between the processes) must reproduce the reference goldens bit for bit -- ghost zones,
EMF correction and the dt reduction all cross rank boundaries.

  RANK=r WORLD_SIZE=n AB_ID_DIR=dir python tests/multirank_emu_check.py golden [golden ...]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    sys.path.insert(0, p)
import gpu_util  # noqa: E402
import util  # noqa: E402


def main(names):
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    iddir = os.environ["AB_ID_DIR"]
    bad = 0
    for name in names:
        g = util.Golden(name)
        if len(g.locs) < world:
            continue
        m = gpu_util.mesh_from_golden(g, rank=rank, nranks=world, device=0)
        path = os.path.join(iddir, "id_" + name)

        def bcast(data, path=path):        # the host application's MPI_Bcast of the NCCL id
            if data is not None:
                with open(path + ".tmp", "wb") as fh:
                    fh.write(data)
                os.replace(path + ".tmp", path)
                return data
            for _ in range(60000):
                if os.path.exists(path):
                    return open(path, "rb").read()
                time.sleep(0.001)
            raise RuntimeError("no NCCL id from rank 0")
        m.init_comm(bcast)
        m.initialize()
        ok = (m.dt == g.dts[0])
        dts = m.cycles(g.ncycles)
        ok &= list(dts) == list(g.dts[:g.ncycles]) and m.dt == g.dts[g.ncycles]
        nbad = 0
        nmine = 0
        for n, loc in enumerate(g.locs):         # loc carries the level on refined meshes
            pmb = m.block_of(*loc)
            if pmb is None:
                continue
            nmine += 1
            for f in g.fields:
                if not np.array_equal(pmb.get(f), g.final[n][f]):
                    nbad += 1
        ok &= (nbad == 0) and nmine == m.nblocal
        # HistoryOutput sums over the blocks of ALL ranks (device reduction + all-reduce SUM) against
        # the reference's .hst row of the final state; summation order differs: 1e-13 of the scale
        if g.hst is not None:
            h = m.history()
            ref = g.hst[g.ncycles, 2:2 + len(h)]
            ok &= bool(np.all(np.abs(h - ref) <= 1e-13*np.abs(ref).max()))
        print("rank %d/%d %s: blocks %d dt_ok %s bad_arrays %d -> %s" %
              (rank, world, name, m.nblocal, list(dts) == list(g.dts[:g.ncycles]), nbad,
               "OK" if ok else "FAIL"), flush=True)
        bad += 0 if ok else 1
        del m
    print("rank %d done: %d failed" % (rank, bad), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
