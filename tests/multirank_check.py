"""Multi-process parity check (one process per GPU, launched by torch.distributed.run):
MeshBlocks are sharded over the ranks exactly like Mesh::CalculateLoadBalance does, ghost zones
and EMFs cross ranks through NCCL, dt through an NCCL MIN all-reduce; every rank compares its
blocks with the reference golden after N cycles (bit-exact)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    sys.path.insert(0, p)
import gpu_util  # noqa: E402
import util  # noqa: E402


DEFAULT_NAMES = ["c5_blast_hlld_plm_vl2_8blk", "c2_linwave_hlld_plm_vl2_8blk",
                 "c4_kh_hllc_ppm_rk2_8blk", "c3_ot_hlld_ppm_vl2_4blk",
                 "c1_sod_hllc_plm_vl2_2blk", "khs3d_mhd_hlld_plm_vl2_8blk_s1",
                 "khs_lhllc_plm_vl2_4blk_s1"]
# statically refined meshes with the levels sharded over the ranks: bit-exact at 2-8 ranks on the
# emulated device path (tests/test_multirank_cpu.py); their first run under NCCL on GPUs was cut
# off by the round's GPU budget before it reported, so they are not part of the default set that
# bench.py runs behind its timed region (python tests/multirank_check.py <names> runs them)
SMR_NAMES = ["smr_blast3d_hllc_plm_vl2", "smr_khs2d_lhllc_plm_vl2_s1",
             "smr_blast2d_lvl2_bcs_hllc_plm_rk2"]


def check_goldens(rank, world, local, names=None, verbose=True):
    """Every golden with at least `world` MeshBlocks, sharded over the ranks of the already
    initialised NCCL process group; returns (fixtures run, fixtures failed on ANY rank).
    bench.py calls this after its timed region so that the scaling record carries multi-rank
    correctness (checker only: nothing here is timed)."""
    nrun = nfail = 0
    for name in (names or DEFAULT_NAMES):
        g = util.Golden(name)
        if len(g.locs) < world:
            continue
        m = gpu_util.mesh_from_golden(g, rank=rank, nranks=world, device=local)

        def bcast(data):
            t = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                t.copy_(torch.tensor(list(data), dtype=torch.uint8))
            dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        m.init_comm(bcast)
        m.initialize()
        good = (m.dt == g.dts[0])
        dts = m.cycles(g.ncycles)
        good &= list(dts) == list(g.dts[:g.ncycles]) and m.dt == g.dts[g.ncycles]
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
        good &= (nbad == 0) and nmine == m.nblocal
        if verbose:
            print("rank %d/%d %s: blocks %d dt_ok %s bad_arrays %d -> %s" %
                  (rank, world, name, m.nblocal, list(dts) == list(g.dts[:g.ncycles]), nbad,
                   "OK" if good else "FAIL"), flush=True)
        t = torch.tensor([0 if good else 1], device="cuda")
        dist.all_reduce(t)
        nrun += 1
        nfail += 1 if int(t.item()) else 0
        del m
    return nrun, nfail


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nrun, nfail = check_goldens(rank, world, local, sys.argv[1:] or None)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if (nfail or not nrun) else 0)


if __name__ == "__main__":
    main()
