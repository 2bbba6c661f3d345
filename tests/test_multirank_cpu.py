"""CPU, world_size 2 and 3, gloo: the host logic of the N>1 path.  MeshBlocks are sharded over
ranks like Mesh::CalculateLoadBalance; every rank derives, independently, the list of messages
it sends to / receives from each peer (ghost zones and EMF correction).  NCCL send/recv pairs
them purely by ORDER and SIZE inside one concatenated buffer per peer, so the test checks that
rank a's send list towards b is exactly rank b's receive list from a -- for periodic and
outflow meshes in 1-D/2-D/3-D -- and that the id broadcast helper works over a process group."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # athinput, overrides, mhd, flux, nghost
    ("athinput.blast", ["mesh/nx1=64", "mesh/nx2=64", "mesh/nx3=64", "meshblock/nx1=16",
                        "meshblock/nx2=32", "meshblock/nx3=32"], True, "hlld", 2),
    ("athinput.orszag_tang", ["mesh/nx1=64", "mesh/nx2=64", "meshblock/nx1=16",
                              "meshblock/nx2=32"], True, "hlld", 3),
    ("athinput.kh", ["mesh/nx1=32", "mesh/nx2=32", "mesh/nx3=32", "meshblock/nx1=16",
                     "meshblock/nx2=16", "meshblock/nx3=16"], False, "hllc", 3),
    ("athinput.sod", ["mesh/nx1=64", "meshblock/nx1=8"], False, "hllc", 2),
    # passive scalars ride behind u (and b) in the ghost-zone messages
    ("athinput.kh_scalar", ["mesh/nx1=32", "mesh/nx2=32", "mesh/nx3=32", "meshblock/nx1=16",
                            "meshblock/nx2=16", "meshblock/nx3=16"], True, "hlld", 2, 2),
]


def reference_load_balance(nb, nranks):
    """Mesh::CalculateLoadBalance with unit costs (src/mesh/amr_loadbalance.cpp:72-112)"""
    rlist = [0]*nb
    total, j = float(nb), nranks - 1
    target, mycost = total/nranks, 0.0
    for i in range(nb - 1, -1, -1):
        mycost += 1.0
        rlist[i] = j
        if mycost >= target and j > 0:
            j -= 1
            total -= mycost
            mycost = 0.0
            target = total/(j + 1)
    return rlist


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import athena_gamma_b200 as ab
    errors = []
    for case in CASES:
        inp, ov, mhd, flux, ng = case[:5]
        ns = case[5] if len(case) > 5 else 0
        pin = ab.ParameterInput(path=os.path.join(ROOT, "inputs", inp))
        pin.modify_from_cmdline(ov)
        plan = ab.MeshPlan(pin, mhd, flux, nghost=ng, rank=rank, nranks=world, nscalars=ns)
        mine = {"ranklist": plan.ranklist(), "nblocal": plan.nblocal,
                "gids": [b.gid for b in plan.my_blocks],
                "msgs": [plan.messages(0), plan.messages(1) if mhd else []]}
        allp = [None]*world
        dist.all_gather_object(allp, mine)
        rl = allp[0]["ranklist"]
        if any(p["ranklist"] != rl for p in allp):
            errors.append("%s: ranks disagree on the load balance" % inp)
        if rl != reference_load_balance(len(rl), world):
            errors.append("%s: load balance differs from CalculateLoadBalance" % inp)
        if sorted(g for p in allp for g in p["gids"]) != list(range(len(rl))):
            errors.append("%s: blocks not partitioned" % inp)
        if mine["gids"] != [g for g, r in enumerate(rl) if r == rank]:
            errors.append("%s: local blocks are not the contiguous gid range" % inp)
        for kind in (0, 1):
            for a in range(world):
                for b in range(world):
                    if a == b:
                        continue
                    snd = [(m["key"], m["count"]) for m in allp[a]["msgs"][kind]
                           if m["dir"] == 0 and m["peer"] == b]
                    rcv = [(m["key"], m["count"]) for m in allp[b]["msgs"][kind]
                           if m["dir"] == 1 and m["peer"] == a]
                    if snd != rcv:
                        errors.append("%s kind %d: send list %d->%d != recv list" % (inp, kind, a, b))
                    if [k for k, _ in snd] != sorted(k for k, _ in snd):
                        errors.append("%s kind %d: messages %d->%d not in key order" % (inp, kind, a, b))
        # every message is addressed to a block the receiver owns
        for m in mine["msgs"][0] + mine["msgs"][1]:
            if m["dir"] == 0 and rl[m["key"]//64] != m["peer"]:
                errors.append("%s: message to gid %d sent to the wrong rank" % (inp, m["key"]//64))

    # the NCCL-id broadcast helper used by Mesh.init_comm, over this process group
    def bcast(data):
        t = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            t.copy_(torch.tensor(list(data), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.tolist())
    payload = bytes(range(128)) if rank == 0 else None
    if bcast(payload) != bytes(range(128)):
        errors.append("id broadcast failed")
    q.put((rank, errors))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_message_plans_pair_up_across_ranks(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, errors in results:
        assert not errors, "rank %d: %s" % (rank, errors[:5])


# ---------------------------------------------------------------------------------------------
# The N > 1 path END TO END on the CPU: world_size processes, each running the emulated device
# path (tests/hostcheck: the product's csrc/*.cu compiled for the host) with the MeshBlocks of a
# golden fixture sharded over the ranks.  The NCCL entry points the library loads with dlopen are
# served by a test-only stand-in over UNIX sockets (tests/hostcheck/nccl_emu.c, AB_NCCL_LIB), so
# the pack / send / receive / unpack plans, the EMF correction across rank boundaries and the dt
# all-reduce really move the data -- and must land on the reference's bits.
SMR_GOLDENS = ["smr_blast2d_hllc_plm_vl2", "smr_blast2d_lvl2_bcs_hllc_plm_rk2",
               "smr_blast3d_hllc_plm_vl2", "smr_blast3d_refl_lhllc_plm_rk3",
               "smr_khs2d_lhllc_plm_vl2_s1", "smr_khs3d_hllc_ppm_rk3_ng4_s2", "smr_sod1d_hllc_plm_vl2"]
EMU_CASES = [
    (2, "0", ["c5_blast_hlld_plm_vl2_8blk", "c2_linwave_hlld_plm_vl2_8blk", "c4_kh_hllc_ppm_rk2_8blk",
              "c3_ot_hlld_ppm_vl2_4blk", "c1_sod_hllc_plm_vl2_2blk", "khs3d_mhd_hlld_plm_vl2_8blk_s1",
              "blast_refl_hlld_plm_vl2_8blk", "iso_blast_hlle_plm_vl2_8blk"]),
    (2, "1", ["c5_blast_hlld_plm_vl2_8blk", "c3_ot_hlld_ppm_vl2_4blk", "blast_mixedbc_hllc_plm_vl2_8blk"]),
    (3, "0", ["c5_blast_hlld_plm_vl2_8blk", "c3_ot_hlld_ppm_vl2_4blk", "iso_blast_hlle_plm_vl2_8blk"]),
    (4, "0", ["c5_blast_hlld_plm_vl2_8blk", "c4_kh_hllc_ppm_rk2_8blk", "c3_ot_hlld_ppm_vl2_4blk",
              "blast_mixedbc_hllc_plm_vl2_8blk"]),
    (8, "0", ["c5_blast_hlld_plm_vl2_8blk", "khs3d_mhd_hlld_plm_vl2_8blk_s1", "blast_hlld_ppm_rk3_8blk"]),
    # statically refined meshes across ranks: ghost zones between levels (restricted slabs,
    # coarse-buffer fills), flux correction fine -> coarse, prolongation next to rank boundaries
    (2, "0", SMR_GOLDENS), (4, "0", SMR_GOLDENS), (8, "0", SMR_GOLDENS[:4]),      # (3 ranks: by hand)
]


@pytest.fixture(scope="module")
def emu_multirank_env(tmp_path_factory):
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "hostcheck"))
    import build_mesh_host
    so = build_mesh_host.build()
    nccl = os.path.join(here, "hostcheck", "libnccl_emu.so")
    src = os.path.join(here, "hostcheck", "nccl_emu.c")
    if not os.path.exists(nccl) or os.path.getmtime(nccl) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O1", "-g", "-fPIC", "-shared", "-o", nccl, src], check=True)
    env = dict(os.environ, AB_LIB=so, AB_NCCL_LIB=nccl, CUDA_VISIBLE_DEVICES="")
    return env


@pytest.mark.parametrize("world,overlap,names", EMU_CASES)
def test_goldens_sharded_over_ranks_through_the_emulated_device_path(emu_multirank_env, tmp_path,
                                                                     world, overlap, names):
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    procs = []
    for r in range(world):
        env = dict(emu_multirank_env, RANK=str(r), WORLD_SIZE=str(world), AB_ID_DIR=str(tmp_path),
                   AB_OVERLAP=overlap)
        procs.append(subprocess.Popen([sys.executable, os.path.join(here, "multirank_emu_check.py")]
                                      + names, env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("rank %d done: 0 failed" % r) in out, out[-3000:]
        assert out.count("-> OK") == len(names), out[-3000:]


@pytest.mark.parametrize("world", [2, 4])
def test_per_block_boundary_tasks_across_ranks_under_a_polling_scheduler(emu_multirank_env, tmp_path,
                                                                         world):
    """tests/sched_check.py with the MeshBlocks sharded over the ranks: every rank runs its own
    randomly drifting TaskList-style scheduler over the C ABI's per-block tasks; ab_bvals_send
    packs for the other ranks while same-rank neighbours may still be a task behind (their
    registers not yet swapped), the grouped NCCL exchange of a round goes out with the last local
    Send, ab_bvals_recv_try / ab_emf_recv_try answer "not yet" until it has.  dt sequence and every
    local array bit-identical to the reference goldens."""
    import subprocess
    import test_gpu_sched
    here = os.path.dirname(os.path.abspath(__file__))
    procs = []
    for r in range(world):
        env = dict(emu_multirank_env, RANK=str(r), WORLD_SIZE=str(world), AB_ID_DIR=str(tmp_path))
        procs.append(subprocess.Popen([sys.executable, os.path.join(here, "sched_check.py"),
                                       "--seed", "3"] + test_gpu_sched.SCHED_GOLDENS, env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=900)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for p, out in zip(procs, outs):
        assert p.returncode == 0 and "sched done: 0 failed" in out, out[-3000:]
