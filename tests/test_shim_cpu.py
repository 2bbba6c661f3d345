"""CPU: the drop-in itself.  The reference is compiled with this repository's shim
(tools/build_shim.py: src/task_list/time_integrator.cpp and src/hydro/new_blockdt.cpp replaced
by shim/*.cpp, everything else -- main(), ParameterInput, Mesh, the C++ problem generators,
Mesh::Initialize, TaskList::DoTaskListOneStage with its OpenMP loop, outputs -- unchanged) and
linked against the C ABI.  Here the library behind the ABI is the emulated device path
(tests/hostcheck, test infrastructure: the product's csrc compiled for the host), so the whole
chain athinput -> pgen -> task list -> ab_* -> kernels -> restart / history files runs without
a GPU and must reproduce the goldens of the UNMODIFIED reference bit for bit.
tests/test_gpu_shim.py runs the same binaries against libathena_b200.so on the GPU."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

# BASELINE.json configs C1-C5 at fixture size, then user-enrolled boundary and source
# functions (the reference's pgen/shk_cloud.cpp, oracle/pgen/usersrc.cpp), passive scalars,
# isothermal EOS, reflecting / mixed boundaries, 1-D, rk3 + PPM, characteristic PLM, gravity
SHIM_GOLDENS = ["c1_sod_hllc_plm_vl2_2blk", "c2_linwave_hlld_plm_vl2_8blk",
                "c3_ot_hlld_ppm_vl2_4blk", "c4_kh_hllc_ppm_rk2_8blk", "c5_blast_hlld_plm_vl2_8blk",
                "shkcloud3d_hlld_plm_vl2_8blk", "usersrc_hlld_plm_rk3_8blk",
                "khs3d_mhd_hlld_plm_vl2_8blk_s1", "iso_blast_hlle_plm_vl2_8blk",
                "blast_refl_hlld_plm_vl2_8blk", "blast_mixedbc_hllc_plm_vl2_8blk",
                "bw1d_hlld_plm_vl2_2blk", "blast_hlld_ppm_rk3_8blk", "blast_hlld_plmc_vl2_8blk",
                "blast_grav_hlld_plm_rk2_8blk"]


@pytest.fixture(scope="module")
def shim_env():
    sys.path.insert(0, os.path.join(HERE, "hostcheck"))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import build_mesh_host
    so = build_mesh_host.build()
    if os.path.isdir("/root/reference/src"):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "build_shim.py")], check=True,
                       capture_output=True)
    libdir = os.path.join(HERE, "hostcheck", "_gen", "emu_lib")
    os.makedirs(libdir, exist_ok=True)
    link = os.path.join(libdir, "libathena_b200.so")
    if os.path.lexists(link):
        os.remove(link)
    os.symlink(so, link)
    env = dict(os.environ)
    env["AB_SHIM_LIBDIR"] = libdir
    env["CUDA_VISIBLE_DEVICES"] = ""
    return env


@pytest.mark.parametrize("threads", [1, 2])
def test_reference_with_shim_reproduces_reference_goldens(shim_env, threads):
    """threads = 2: the reference's OpenMP loop over MeshBlocks (task_list.cpp:71-88) calls the
    per-block entry points of the C ABI from two host threads"""
    names = SHIM_GOLDENS if threads == 2 else SHIM_GOLDENS[:5]    # bounds the suite's run time
    chunks = [names[c::3] for c in range(3)]
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "shim_check.py"), "--threads",
                               str(threads)] + ch, env=shim_env, stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for ch in chunks]
    for ch, p in zip(chunks, procs):
        out, err = p.communicate(timeout=1500)
        assert p.returncode == 0 and "shim done: 0 failed" in out, out[-3000:] + err[-2000:]
        assert out.count("\nok ") + out.startswith("ok ") == len(ch)
