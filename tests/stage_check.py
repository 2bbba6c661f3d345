"""Run by tests/test_gpu_staging.py in its own process: the pipelined staging entry points
(ab_stage_*) must give, for every step, exactly what the plain ab_upload -> initialize ->
cycle -> ab_download sequence gives, while uploads / downloads overlap the kernels."""
import ctypes as C
import os
import sys

import numpy as np
import torch


def _pin(t):
    """pinned where there is a driver (the emulated library of test_emu_device_path has none)"""
    return t.pin_memory() if torch.cuda.is_available() else t


HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    sys.path.insert(0, p)
import athena_gamma_b200 as ab  # noqa: E402
import gpu_util  # noqa: E402
import util  # noqa: E402


def main():
    g = util.Golden("c5_blast_hlld_plm_vl2_8blk")
    m = gpu_util.mesh_from_golden(g)
    L, dp = m.L, C.POINTER(C.c_double)
    names = list(g.fields)
    nsteps = 4
    # step s starts from the golden initial state scaled by (1 + 0.01 s): distinct inputs per step
    inputs = []
    for s in range(nsteps):
        inputs.append([{f: _pin(torch.from_numpy(np.ascontiguousarray(
            g.init[n][f]*(1.0 + 0.01*s)))) for f in names}
            for n in range(len(g.locs))])
    order = [m.block_of(*loc) for loc in g.locs]

    L.ab_mesh_set_time_dt.argtypes = [C.c_void_p, C.c_double, C.c_double]

    def reset_clock():
        # same (time, dt_old) in front of every step of both sequences: Mesh::NewTimeStep limits
        # the new dt by 2*dt_old (mesh/mesh.cpp:1078-1119)
        ab.lib.check(L.ab_mesh_set_time_dt(m.h, 0.0, float(np.finfo(np.float64).max)))

    def plain(s):
        for pmb, d in zip(order, inputs[s]):
            for f in names:
                pmb.set(f, d[f].numpy())
        reset_clock()
        m.initialize()
        m.cycles(1)
        return [{f: pmb.get(f).copy() for f in names} for pmb in order]
    want = [plain(s) for s in range(nsteps)]

    nreg = len(names)
    regs = (C.c_int*nreg)(*[ab.lib.REG[f] for f in names])
    ab.lib.check(L.ab_stage_begin(m.h, regs, nreg))
    by_lid = sorted(range(len(order)), key=lambda n: order[n].lid)
    outs = [[{f: _pin(torch.zeros_like(inputs[0][n][f])) for f in names}
             for n in range(len(order))] for _ in range(nsteps)]

    def ptrs(bufs):
        return (dp*(nreg*len(order)))(*[C.cast(bufs[n][f].data_ptr(), dp)
                                        for n in by_lid for f in names])
    ab.lib.check(L.ab_stage_upload_all(m.h, ptrs(inputs[0])))
    for s in range(nsteps):
        ab.lib.check(L.ab_stage_commit(m.h))
        if s + 1 < nsteps:
            ab.lib.check(L.ab_stage_upload_all(m.h, ptrs(inputs[s + 1])))
        reset_clock()
        ab.lib.check(L.ab_mesh_initialize(m.h))
        ab.lib.check(L.ab_mesh_cycles(m.h, 1))
        ab.lib.check(L.ab_stage_download_all(m.h, ptrs(outs[s])))
    ab.lib.check(L.ab_stage_sync(m.h))
    for s in range(nsteps):
        for n in range(len(order)):
            for f in names:
                util.assert_bitwise(outs[s][n][f].numpy(), want[s][n][f],
                                    "step %d block %d %s" % (s, n, f))
    print("staging ok")
    return 0


if __name__ == "__main__":
    sys.exit(main())
