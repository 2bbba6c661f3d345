"""Run by tests/test_shim_cpu.py (emulated library, CPU) and tests/test_gpu_shim.py (GPU): the
reference compiled WITH the shim (tools/build_shim.py: its own main(), ParameterInput, Mesh,
C++ problem generator, Mesh::Initialize, polling task scheduler and outputs; task bodies = the C
ABI of libathena_b200) re-runs the parameter blocks of a golden fixture -- which the UNMODIFIED
reference binary produced -- with a restart dump and a history line every cycle, and must
reproduce it bit for bit: the dt sequence of the stdout cycle lines, u / b / s of every
MeshBlock including ghost zones in the final restart file, every column of the .hst rows.

  python tests/shim_check.py [--threads N] golden [golden ...]
env AB_SHIM_LIBDIR: directory holding the libathena_b200.so to run against (the emulated
library for the CPU test); default = the binary's RUNPATH (athena-gamma_b200/).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools"), HERE):
    sys.path.insert(0, p)
import ref_run  # noqa: E402
import util  # noqa: E402

SHIM_BUILD = os.path.join(ROOT, "shim", "_build")


def shim_exe(cfg, pgen):
    return os.path.join(SHIM_BUILD, cfg, "athena_" + pgen)


def run_golden(name, threads):
    g = util.Golden(name)
    cfg, pgen = g.meta["cfg"], g.meta["pgen"]
    exe = shim_exe(cfg, pgen)
    if not os.path.isfile(exe):
        raise FileNotFoundError(exe + " (python tools/build_shim.py)")
    blocks = {b: dict(kv) for b, kv in g.par.items()}
    for b, kv in blocks.items():            # state the reference wrote back into its dump
        if b.startswith("output"):
            kv.pop("next_time", None)
            kv.pop("file_number", None)
    env = {}
    if os.environ.get("AB_SHIM_LIBDIR"):
        env["LD_LIBRARY_PATH"] = os.environ["AB_SHIM_LIBDIR"] + ":" + os.environ.get("LD_LIBRARY_PATH", "")
    res = ref_run.run_reference(cfg, pgen, None, None, threads=threads, exe=exe, env_extra=env,
                                blocks=blocks, hst_every_cycle=False, timeout=900)
    try:
        assert "[b200] device mesh created" in res["stdout"], "the shim did not engage"
        n = g.ncycles
        dts = res["dts"][:n + 1]
        assert list(dts) == list(g.dts[:n + 1]), "dt sequence: %r vs %r" % (dts, list(g.dts))
        nhydro = 4 if g.eos == "isothermal" else 5
        last = ref_run.read_rst(res["rst"][n], nhydro=nhydro, mhd=g.mhd, nscalars=g.nscalars)
        assert last["time"] == g.final_time and last["dt"] == g.final_dt
        locs = [tuple(b["loc"][:3]) for b in last["blocks"]]
        for k, loc in enumerate(g.locs):
            b = last["blocks"][locs.index(tuple(loc[:3]))]
            for f in g.fields:
                util.assert_bitwise(b[f], g.final[k][f], "%s %s block %d" % (name, f, k))
        if g.hst is not None:
            pid = blocks["job"]["problem_id"]
            rows = np.array([[float(x) for x in ln.split()]
                             for ln in open(os.path.join(res["dir"], pid + ".hst"))
                             if not ln.startswith("#")])
            util.assert_bitwise(rows[:n + 1], g.hst[:n + 1], name + " history rows")
    finally:
        ref_run.cleanup(res)


def main():
    args = sys.argv[1:]
    threads = None
    if args and args[0] == "--threads":
        threads = int(args[1])
        args = args[2:]
    bad = 0
    for name in args:
        try:
            run_golden(name, threads)
            print("ok %s" % name, flush=True)
        except Exception as ex:   # noqa: BLE001
            bad += 1
            print("FAILED %s: %s" % (name, str(ex)[:1500]), flush=True)
    print("shim done: %d failed" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
