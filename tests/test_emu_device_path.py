"""CPU: the product's whole device path -- kernels, launch geometry, work lists, the mesh / task
glue and the C ABI of athena-gamma_b200/csrc -- compiled for the host against a minimal CUDA
emulation (tests/hostcheck/emu/cuda_runtime.h: grids run block by block, thread by thread; the
runtime API synchronous, device memory NaN-filled on allocation) and driven through the same
Python Mesh as the GPU tests, against the reference's golden vectors, bit for bit.

What this covers without a GPU: every golden fixture (dt sequence + final state of every
MeshBlock), the statically refined meshes (ab_mesh_create_refined and the SMR launch glue) and
the ab_stage_* pipeline.  What it cannot cover: anything that depends on parallel execution
(races between threads or streams), NCCL, and the numerics of nvcc's code generation -- those
stay with the -m gpu tests.  The emulated library is test infrastructure only; the product
loads libathena_b200.so and fails loudly without a CUDA device (tests/test_abi.py)."""
import os
import subprocess
import sys

import pytest

import util

HERE = os.path.dirname(os.path.abspath(__file__))
NCHUNK = 4


@pytest.fixture(scope="module")
def emu_env():
    sys.path.insert(0, os.path.join(HERE, "hostcheck"))
    import build_mesh_host
    so = build_mesh_host.build()
    env = dict(os.environ)
    env["AB_LIB"] = so
    env["CUDA_VISIBLE_DEVICES"] = ""
    return env


def in_device_scope(name):
    if not name.startswith("smr_"):
        return True
    import test_gpu_smr
    return name in test_gpu_smr.device_smr_goldens()


@pytest.mark.parametrize("order", ["forward", "reverse"])
def test_every_golden_through_the_emulated_device_path(emu_env, order):
    """order = reverse runs the blocks of a grid and the threads of a block last to first: a
    kernel in which one thread reads what another thread of the same launch writes (a race on
    the GPU) gives a different answer in the two orders"""
    emu_env = dict(emu_env, AB_EMU_ORDER=order)
    names = [n for n in util.golden_names(include_smr=True) if in_device_scope(n)]
    assert len(names) >= 60
    chunks = [names[c::NCHUNK] for c in range(NCHUNK)]
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "smr_check.py")] + ch,
                              env=emu_env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for ch in chunks]
    bad = []
    for ch, p in zip(chunks, procs):
        try:
            out = p.communicate(timeout=600)[0]
        except subprocess.TimeoutExpired:
            p.kill()
            out = "TIMEOUT " + " ".join(ch)
        if p.returncode != 0 or "smr done: 0 failed" not in out:
            bad.append(out[-3000:])
        else:
            assert out.count("\nok ") + out.startswith("ok ") == len(ch)
    assert not bad, "\n".join(bad)


SCHEDULE_GOLDENS = ["c5_blast_hlld_plm_vl2_8blk", "c3_ot_hlld_ppm_vl2_4blk",
                    "c4_kh_hllc_ppm_rk2_8blk", "blast_mixedbc_hllc_plm_vl2_8blk",
                    "khs3d_mhd_hlld_plm_vl2_8blk_s1", "c1_sod_hllc_plm_vl2_2blk"]


@pytest.mark.parametrize("knob", ["AB_NO_BATCH", "AB_OVERLAP"])
def test_other_schedules_through_the_emulated_device_path(emu_env, knob):
    """The cycle runs every task as ONE launch over all MeshBlocks of the rank (ab_batch.cuh) by
    default; AB_NO_BATCH=1 launches block by block, AB_OVERLAP=1 is the multi-GPU schedule with
    the Primitives task split into active cells and a one-launch ghost shell.  All three must
    land on the same bits (periodic, mixed physical boundaries, scalars, 1-D / 2-D / 3-D)."""
    env = dict(emu_env)
    env[knob] = "1"
    r = subprocess.run([sys.executable, os.path.join(HERE, "smr_check.py")] + SCHEDULE_GOLDENS,
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "smr done: 0 failed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_staged_pipeline_through_the_emulated_device_path(emu_env):
    r = subprocess.run([sys.executable, os.path.join(HERE, "stage_check.py")], env=emu_env,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "staging ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_polling_scheduler_through_the_emulated_device_path(emu_env):
    """tests/sched_check.py: the per-block boundary tasks under a host scheduler that polls the
    way TaskList::DoTaskListOneStage does, blocks drifting apart by whole tasks"""
    import test_gpu_sched
    names = test_gpu_sched.SCHED_GOLDENS
    chunks = [names[c::3] for c in range(3)]          # three processes side by side
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "sched_check.py"), "--seed", "3"]
                              + ch, env=emu_env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for ch in chunks]
    for p in procs:
        out = p.communicate(timeout=900)[0]
        assert p.returncode == 0 and "sched done: 0 failed" in out, out[-3000:]


def test_every_task_entry_point_through_the_emulated_device_path(emu_env):
    """tests/test_gpu_tasks.py (each C-ABI task against the oracle on perturbed states that hit
    floors, limiter and solver branches) with the emulated library"""
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_gpu_tasks.py"),
                        "-q", "-m", "gpu", "-p", "no:cacheprovider", "--timeout",
                        "300", "-n", "4"], env=emu_env, capture_output=True, text=True,
                       timeout=900)
    tail = r.stdout.strip().splitlines()[-1]
    assert r.returncode == 0 and " passed" in tail and "failed" not in tail, r.stdout[-3000:]


def test_no_out_of_bounds_access_under_address_sanitizer(emu_env):
    """the same path with AddressSanitizer: "device" allocations are heap blocks with red zones,
    so an out-of-range index in any kernel or copy aborts the run.  A fifth of the fixtures, the
    PPM fixtures (shared-memory / shuffle kernels) and three refined meshes (all fixtures pass
    when run by hand; the subset bounds the suite's run time)."""
    import build_mesh_host
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True,
                             text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("libasan not installed")
    env = dict(emu_env, AB_LIB=build_mesh_host.build(asan=True), LD_PRELOAD=libasan,
               ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    names = [n for n in util.golden_names(include_smr=True) if in_device_scope(n)]
    smr = [n for n in names if n.startswith("smr_")][:3]
    names = [n for i, n in enumerate(names) if not n.startswith("smr_") and
             (i % 5 == 0 or n in ("c3_ot_hlld_ppm_vl2_4blk", "c4_kh_hllc_ppm_rk2_8blk"))] + smr
    chunks = [names[c::NCHUNK] for c in range(NCHUNK)]
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "smr_check.py")] + ch, env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for ch in chunks]
    for ch, p in zip(chunks, procs):
        try:
            out = p.communicate(timeout=900)[0]
        except subprocess.TimeoutExpired:
            p.kill()
            out = "TIMEOUT " + " ".join(ch)
        assert p.returncode == 0 and "smr done: 0 failed" in out, out[-4000:]
        assert "ERROR: AddressSanitizer" not in out, out[-4000:]


def test_emulator_positive_controls(tmp_path):
    """the checks above can fail: a kernel with a dependence between threads of one launch
    differs between the two execution orders, an out-of-range store aborts under
    AddressSanitizer, and the fibers give the right block sums for shuffles + __syncthreads"""
    hc = os.path.join(HERE, "hostcheck")
    exe, exe_asan = str(tmp_path / "selftest"), str(tmp_path / "selftest_asan")
    base = ["g++", "-O1", "-std=c++17", "-I", os.path.join(hc, "emu"),
            os.path.join(hc, "emu_selftest.cpp")]
    subprocess.run(base + ["-o", exe], check=True)

    def run(mode, order, binary=exe):
        return subprocess.run([binary, mode], env=dict(os.environ, AB_EMU_ORDER=order),
                              capture_output=True, text=True)
    assert run("shift", "forward").stdout != run("shift", "reverse").stdout
    sums = "8128 24512 40896 57280\n"
    assert run("reduce", "forward").stdout == sums and run("reduce", "reverse").stdout == sums
    if subprocess.run(base + ["-fsanitize=address", "-o", exe_asan]).returncode == 0:
        r = run("oob", "forward", exe_asan)
        assert r.returncode != 0 and "heap-buffer-overflow" in r.stderr
