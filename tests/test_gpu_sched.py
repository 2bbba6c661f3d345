"""GPU: the per-MeshBlock boundary tasks of the C ABI (ab_bvals_send / ab_bvals_recv_try /
ab_bvals_set, ab_emf_send / ab_emf_recv_try, ab_clear_boundary) under a polling host scheduler
that mirrors TaskList::DoTaskListOneStage (task_list/task_list.cpp:66-91) with the blocks
drifting apart by whole tasks -- see tests/sched_check.py.  In its own process so that a CUDA
error there cannot poison the context of the other tests."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

# every boundary kind, integrator and variable set the per-block tasks distinguish
SCHED_GOLDENS = ["c5_blast_hlld_plm_vl2_8blk", "c4_kh_hllc_ppm_rk2_8blk", "c3_ot_hlld_ppm_vl2_4blk",
                 "c2_linwave_hlld_plm_vl2_8blk", "c1_sod_hllc_plm_vl2_2blk",
                 "khs3d_mhd_hlld_plm_vl2_8blk_s1", "usersrc_hlld_plm_rk3_8blk",
                 "shkcloud3d_hlld_plm_vl2_8blk", "blast_refl_hlld_plm_vl2_8blk",
                 "blast_mixedbc_hllc_plm_vl2_8blk", "iso_ot_hlld_plm_rk2_4blk",
                 "blast_noncubic_hlld_ppm_rk2_6blk", "bw1d_hlld_plm_vl2_2blk"]


@pytest.mark.parametrize("seed", [1, 2])
def test_polling_scheduler_with_per_block_boundary_tasks(seed):
    r = subprocess.run([sys.executable, os.path.join(HERE, "sched_check.py"), "--seed", str(seed)]
                       + SCHED_GOLDENS, capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:], r.stderr[-3000:])
    assert r.returncode == 0 and "sched done: 0 failed" in r.stdout
    assert "answered not-yet" in r.stdout
