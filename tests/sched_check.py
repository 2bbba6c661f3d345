"""Run by tests/test_gpu_sched.py (GPU) and tests/test_emu_device_path.py (emulated library) in
its own process: golden fixtures stepped through the C ABI the way the reference's host code
drives a cycle -- TaskList::DoTaskListOneStage (task_list/task_list.cpp:66-91) polling every
MeshBlock, each block advancing through the TimeIntegratorTaskList task graph
(time_integrator.cpp:899-1098) as far as its dependencies allow, the boundary tasks being the
per-block ab_bvals_send / ab_bvals_recv_try / ab_bvals_set and ab_emf_send / ab_emf_recv_try.
Blocks are visited in a shuffled order every sweep, a block is skipped at random and runs a random
number (1-4) of available tasks per visit, so blocks drift apart by whole tasks (a block can be at CONS2PRIM while its neighbour
has not integrated yet): the schedule-dependent interleavings a host scheduler can produce.
dt comes from ab_new_block_dt per block + Mesh::NewTimeStep on the host (mesh.cpp:1078-1119).
Result: dt sequence and every array of every block bit-identical to the reference golden.

  python tests/sched_check.py [--seed N] golden [golden ...]

With RANK / WORLD_SIZE / AB_ID_DIR in the environment (tests/test_multirank_cpu.py: several
processes of the emulated device path with the socket stand-in for NCCL) the MeshBlocks are
sharded over the ranks: every rank runs its own randomly drifting scheduler, the per-block Send /
ReceiveTry / Set then pack, count rounds and unpack across rank boundaries, and the host's
MPI_Allreduce(MIN) of Mesh::NewTimeStep is a file exchange in AB_ID_DIR.
"""
import ctypes as C
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    sys.path.insert(0, p)
import athena_gamma_b200 as ab  # noqa: E402
import gpu_util  # noqa: E402
import util  # noqa: E402

HYD, FLD, SCL = 0, 1, 2
U, U1, BX1, B1X1, S, S1 = 0, 1, 4, 7, 25, 26

# stage weights (delta, gamma_1, gamma_2, gamma_3, beta): time_integrator.cpp:199-235 (vl2),
# :237-262 (rk1), :425-458 (rk2), :462-520 (rk3)
WEIGHTS = {
    "vl2": [(1.0, 0.0, 1.0, 0.0, 0.5), (0.0, 0.0, 1.0, 0.0, 1.0)],
    "rk1": [(1.0, 0.0, 1.0, 0.0, 1.0)],
    "rk2": [(1.0, 0.0, 1.0, 0.0, 1.0), (0.0, 0.5, 0.5, 0.0, 0.5)],
    "rk3": [(1.0, 0.0, 1.0, 0.0, 1.0), (0.0, 0.25, 0.75, 0.0, 0.25),
            (0.0, 2.0/3.0, 1.0/3.0, 0.0, 2.0/3.0)],
}
# start / end time of each stage in units of dt (IntegratorWeight::sbeta, ebeta)
SBETA = {"vl2": [0.0, 0.5], "rk1": [0.0], "rk2": [0.0, 1.0], "rk3": [0.0, 1.0, 0.5]}
EBETA = {"vl2": [0.5, 1.0], "rk1": [1.0], "rk2": [1.0, 1.0], "rk3": [1.0, 0.5, 1.0]}


def build_tasks(mhd, ns):
    """(name, dependencies) in list order, as the reference's constructor adds them for a
    one-level mesh without STS / orbital advection / shearing box"""
    t = [("CALC_HYDFLX", [])]
    if ns:
        t.append(("CALC_SCLRFLX", ["CALC_HYDFLX"]))
    t.append(("INT_HYD", ["CALC_HYDFLX"]))
    t.append(("SRC_TERM", ["INT_HYD"] + (["INT_SCLR"] if ns else [])))
    t += [("SEND_HYD", ["SRC_TERM"]), ("RECV_HYD", []), ("SETB_HYD", ["RECV_HYD", "SRC_TERM"])]
    if ns:
        t += [("INT_SCLR", ["CALC_SCLRFLX"]), ("SEND_SCLR", ["SRC_TERM"]), ("RECV_SCLR", []),
              ("SETB_SCLR", ["RECV_SCLR", "SRC_TERM"])]
    if mhd:
        t += [("CALC_FLDFLX", ["CALC_HYDFLX"]), ("SEND_FLDFLX", ["CALC_FLDFLX"]),
              ("RECV_FLDFLX", ["SEND_FLDFLX"]), ("INT_FLD", ["RECV_FLDFLX"]),
              ("SEND_FLD", ["INT_FLD"]), ("RECV_FLD", []), ("SETB_FLD", ["RECV_FLD", "INT_FLD"])]
    dep = ["SETB_HYD"] + (["SETB_FLD"] if mhd else []) + (["SETB_SCLR"] if ns else [])
    t += [("CONS2PRIM", dep), ("PHY_BVAL", ["CONS2PRIM"]), ("USERWORK", ["PHY_BVAL"]),
          ("NEW_DT", ["USERWORK"]), ("CLEAR_ALLBND", ["NEW_DT"])]
    return t


RANK, WORLD = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def wait_for(path):
    import time as _t
    for _ in range(120000):
        if os.path.exists(path):
            return open(path, "rb").read()
        _t.sleep(0.0005)
    raise RuntimeError("timeout waiting for " + path)


def put(path, data):
    with open(path + ".tmp%d" % RANK, "wb") as fh:
        fh.write(data)
    os.replace(path + ".tmp%d" % RANK, path)


def allreduce_min(tag, value):
    """the host application's MPI_Allreduce(MIN) (mesh.cpp:1105-1108)"""
    if WORLD == 1:
        return value
    d = os.environ["AB_ID_DIR"]
    put(os.path.join(d, "%s_%d" % (tag, RANK)), repr(value).encode())
    return min(float(wait_for(os.path.join(d, "%s_%d" % (tag, r))).decode()) for r in range(WORLD))


def run_golden(name, seed):
    g = util.Golden(name)
    if WORLD > 1:
        m = gpu_util.mesh_from_golden(g, rank=RANK, nranks=WORLD, device=0)
        idpath = os.path.join(os.environ["AB_ID_DIR"], "schedid_" + name)
        m.init_comm(lambda data: (put(idpath, data), data)[1] if data is not None else wait_for(idpath))
    else:
        m = gpu_util.mesh_from_golden(g)
    m.initialize()
    L, h = m.L, m.h
    ck = ab.lib.check
    integ = g.par["time"].get("integrator", "vl2")
    xorder = m.params.xorder
    wts = WEIGHTS[integ]
    tasks = build_tasks(g.mhd, g.nscalars)
    rng = random.Random(seed + 7919*RANK)     # every rank drifts on its own
    time, dt = float(g.par["time"].get("start_time", 0.0)), m.dt
    tlim = float(g.par["time"]["tlim"])
    assert dt == g.dts[0], (dt, g.dts[0])
    new_dt = {}
    polls = fails = 0

    def wave(out, inp, w0, w1, lid):
        w = (C.c_double*5)(w0, w1, 0.0, 0.0, 0.0)
        ck(L.ab_weighted_ave(h, lid, out, inp, w))

    def integrate(lid, stage, regs, add):
        """IntegrateHydro / Field / Scalars (time_integrator.cpp:1563-1650, 2141-2185)"""
        delta, g1, g2, g3, beta = wts[stage-1]
        r, r1 = regs
        wave(r1, r, 1.0, delta, lid)
        if g1 == 0.0 and g2 == 1.0 and g3 == 0.0:
            ck(L.ab_swap(h, lid, r))
        else:
            wave(r, r1, g1, g2, lid)
        ck(add(h, lid, beta*dt))

    def do_task(tname, pmb, stage):
        nonlocal polls, fails
        lid = pmb.lid
        order = 1 if (integ == "vl2" and stage == 1) else xorder
        if tname == "CALC_HYDFLX":
            ck(L.ab_calc_fluxes(h, lid, order, dt))
        elif tname == "CALC_SCLRFLX":
            ck(L.ab_calc_scalar_fluxes(h, lid, order))
        elif tname == "CALC_FLDFLX":
            ck(L.ab_corner_e(h, lid))
        elif tname == "SEND_FLDFLX":
            ck(L.ab_emf_send(h, lid))
        elif tname == "RECV_FLDFLX":
            polls += 1
            if not ck(L.ab_emf_recv_try(h, lid)):
                fails += 1
                return False
        elif tname == "INT_HYD":
            integrate(lid, stage, (U, U1), L.ab_add_flux_div)
        elif tname == "INT_FLD":
            integrate(lid, stage, (BX1, B1X1), L.ab_ct)
        elif tname == "INT_SCLR":
            integrate(lid, stage, (S, S1), L.ab_add_scalar_flux_div)
        elif tname == "SRC_TERM":
            ck(L.ab_add_source_terms(h, lid, time + SBETA[integ][stage-1]*dt, wts[stage-1][4]*dt))
        elif tname.startswith("SEND_"):
            ck(L.ab_bvals_send(h, lid, {"HYD": HYD, "FLD": FLD, "SCLR": SCL}[tname[5:]]))
        elif tname.startswith("RECV_"):
            polls += 1
            if not ck(L.ab_bvals_recv_try(h, lid, {"HYD": HYD, "FLD": FLD, "SCLR": SCL}[tname[5:]])):
                fails += 1
                return False
        elif tname.startswith("SETB_"):
            ck(L.ab_bvals_set(h, lid, {"HYD": HYD, "FLD": FLD, "SCLR": SCL}[tname[5:]]))
        elif tname == "CONS2PRIM":
            ck(L.ab_primitives(h, lid))
        elif tname == "PHY_BVAL":
            # PhysicalBoundary: t_end_stage, beta*dt (time_integrator.cpp:2045-2062)
            ck(L.ab_physical_bcs_at(h, lid, time + EBETA[integ][stage-1]*dt, wts[stage-1][4]*dt))
        elif tname == "NEW_DT":
            if stage == len(wts):
                d = C.c_double()
                ck(L.ab_new_block_dt(h, lid, C.byref(d)))
                new_dt[lid] = d.value
        elif tname == "CLEAR_ALLBND":
            ck(L.ab_clear_boundary(h, lid))
        return True

    dts = []
    for _ in range(g.ncycles):
        dts.append(dt)
        for stage in range(1, len(wts) + 1):
            # StartupTaskList (time_integrator.cpp:1384-1410)
            if stage == 1:
                for pmb in m.my_blocks:
                    ck(L.ab_zero(h, pmb.lid, U1))
                    if g.mhd:
                        ck(L.ab_zero(h, pmb.lid, B1X1))
                    if g.nscalars:
                        ck(L.ab_zero(h, pmb.lid, S1))
            done = {pmb.lid: set() for pmb in m.my_blocks}
            left = len(m.my_blocks)
            while left:
                blocks = list(m.my_blocks)
                rng.shuffle(blocks)
                for pmb in blocks:
                    fin = done[pmb.lid]
                    if len(fin) == len(tasks) or rng.random() < 0.35:
                        continue         # some blocks are not visited this sweep: blocks drift apart
                    budget = rng.randint(1, 4)
                    # DoAllAvailableTasks: unfinished tasks whose dependencies are clear, in list
                    # order; a task that answers "not yet" is skipped and polled again later
                    for tname, deps in tasks:
                        if tname in fin or not all(d in fin for d in deps):
                            continue
                        if do_task(tname, pmb, stage):
                            fin.add(tname)
                            if len(fin) == len(tasks):
                                left -= 1
                            budget -= 1
                            if budget == 0:
                                break    # TaskStatus::success: on to the next MeshBlock
        time += dt
        # Mesh::NewTimeStep (mesh.cpp:1078-1119)
        dt_new = min(2.0*dt, allreduce_min("dt_%s_%d" % (name, len(dts)), min(new_dt.values())))
        if time < tlim and (tlim - time) < dt_new:
            dt_new = tlim - time
        dt = dt_new
        ck(L.ab_mesh_set_time_dt(h, time, dt))
    assert list(dts) == list(g.dts[:len(dts)]), "dt sequence differs: %r vs %r" % (dts, list(g.dts))
    assert dt == g.final_dt, (dt, g.final_dt)
    nmine = 0
    for n, loc in enumerate(g.locs):
        pmb = m.block_of(*loc)
        if pmb is None:          # a block of another rank
            continue
        nmine += 1
        for f in g.fields:
            util.assert_bitwise(pmb.get(f), g.final[n][f], "%s %s block %d" % (name, f, n))
    assert nmine == m.nblocal
    return polls, fails


def main():
    args = sys.argv[1:]
    seed = 1
    if args and args[0] == "--seed":
        seed = int(args[1])
        args = args[2:]
    bad = 0
    for name in args:
        if WORLD > len(util.Golden(name).locs):
            print("ok %s (skipped: fewer MeshBlocks than ranks)" % name, flush=True)
            continue
        try:
            polls, fails = run_golden(name, seed)
            print("ok %s (%d polls, %d answered not-yet)" % (name, polls, fails), flush=True)
        except Exception as ex:   # noqa: BLE001
            bad += 1
            print("FAILED %s: %s" % (name, str(ex)[:600]), flush=True)
    print("sched done: %d failed" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
