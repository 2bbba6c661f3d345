#!/usr/bin/env python3
"""Run the UNMODIFIED reference binary built by oracle/build_ref.py and read its outputs.

TEST INFRASTRUCTURE ONLY (tests/, golden generation, bench.py's reference arm).

 * `run_reference(cfg, pgen, athinput, overrides, rst_every_cycle=...)` runs
   `oracle/_ref/<cfg>/athena_<pgen>` in a scratch directory and returns the dt sequence
   parsed from the 17-digit `cycle= time= dt=` stdout lines (reference src/mesh/mesh.cpp:1951-1985),
   the zone-cycles/s figures (src/main.cpp:578-596) and the per-cycle restart dumps.
 * `read_rst(path)` parses a restart file (layout: src/outputs/restart.cpp:31-207): parameter
   dump text, header (nbtotal, root_level, RegionSize, time, dt, ncycle, datasize), the
   LogicalLocation+cost list, then per block the raw `u` (NHYDRO,nc3,nc2,nc1 incl. ghosts) and
   `b.x1f, b.x2f, b.x3f`.
"""
import os
import re
import shutil
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def ref_binary(cfg, pgen):
    return os.path.join(REF, cfg, "athena_" + pgen)


def have_ref(cfg, pgen):
    return os.path.isfile(ref_binary(cfg, pgen))


def parse_athinput(text):
    """-> ordered dict block -> ordered dict key -> str (comments stripped)."""
    blocks = {}
    cur = None
    for line in text.splitlines():
        line = line.split("#", 1)[0].strip()
        if not line:
            continue
        m = re.match(r"<(\w+)>", line)
        if m:
            cur = m.group(1)
            blocks.setdefault(cur, {})
            continue
        if "=" in line and cur is not None:
            k, v = line.split("=", 1)
            blocks[cur][k.strip()] = v.strip()
    return blocks


def dump_athinput(blocks):
    out = []
    for b, kv in blocks.items():
        out.append("<%s>" % b)
        for k, v in kv.items():
            out.append("%s = %s" % (k, v))
        out.append("")
    return "\n".join(out)


def apply_overrides(blocks, overrides):
    """overrides: dict 'block/key' -> value (creates blocks/keys as needed)."""
    for bk, v in (overrides or {}).items():
        b, k = bk.split("/", 1)
        blocks.setdefault(b, {})[k] = str(v)
    return blocks


_CYCLE_RE = re.compile(r"cycle=(\d+)\s+time=([-+0-9.eE]+)\s+dt=([-+0-9.eE]+)")


def read_rst(path, nhydro=5, mhd=None, nghost=None, nscalars=0):
    raw = open(path, "rb").read()
    tag = b"<par_end>\n"
    pe = raw.index(tag) + len(tag)
    par = parse_athinput(raw[:pe].decode("ascii", "replace"))
    off = pe
    nbtotal, root_level = struct.unpack_from("<ii", raw, off)
    off += 8
    # RegionSize: 9 Reals (x1min,x2min,x3min,x1max,x2max,x3max,x1rat,x2rat,x3rat) + 3 ints
    # (src/mesh/mesh.hpp RegionSize), padded to 8-byte alignment -> 72 + 12 (+4 pad) = 88
    rs = struct.unpack_from("<9d3i", raw, off)
    off += 88
    time, dt = struct.unpack_from("<dd", raw, off)
    off += 16
    (ncycle,) = struct.unpack_from("<i", raw, off)
    off += 4
    (datasize,) = struct.unpack_from("<Q", raw, off)
    off += 8
    locs = []
    for _ in range(nbtotal):
        lx1, lx2, lx3, lev = struct.unpack_from("<qqqi", raw, off)
        off += 32  # LogicalLocation: 3*int64 + int (+4 pad)
        (cost,) = struct.unpack_from("<d", raw, off)
        off += 8
        locs.append((lx1, lx2, lx3, lev))
    mb = par["meshblock"] if "meshblock" in par else par["mesh"]
    nx = [int(mb.get("nx%d" % d, par["mesh"]["nx%d" % d])) for d in (1, 2, 3)]
    ncell_u = datasize // 8
    # infer nghost / mhd from datasize if not given
    cands = []
    for ng in ((nghost,) if nghost else (2, 3, 4)):
        nc = [n + 2 * ng if n > 1 else 1 for n in nx]
        ncc = nc[0] * nc[1] * nc[2]
        nfc = ((nc[0] + 1) * nc[1] * nc[2] + nc[0] * (nc[1] + 1) * nc[2]
               + nc[0] * nc[1] * (nc[2] + 1))
        for m in ((mhd,) if mhd is not None else (False, True)):
            if (nhydro + nscalars) * ncc + (nfc if m else 0) == ncell_u:
                cands.append((ng, m, nc))
    assert len(cands) == 1, "cannot infer block layout from datasize %d" % datasize
    ng, m, nc = cands[0]
    blocks = []
    for b in range(nbtotal):
        o = off + b * datasize
        u = np.frombuffer(raw, "<f8", nhydro * nc[2] * nc[1] * nc[0], o).reshape(
            nhydro, nc[2], nc[1], nc[0]).copy()
        o += u.nbytes
        blk = {"loc": locs[b], "u": u}
        if m:
            shp = [(nc[2], nc[1], nc[0] + 1), (nc[2], nc[1] + 1, nc[0]),
                   (nc[2] + 1, nc[1], nc[0])]
            for name, s in zip(("b1", "b2", "b3"), shp):
                a = np.frombuffer(raw, "<f8", s[0] * s[1] * s[2], o).reshape(s).copy()
                o += a.nbytes
                blk[name] = a
        if nscalars > 0:      # outputs/restart.cpp:172-178: s follows u (and b)
            a = np.frombuffer(raw, "<f8", nscalars * nc[2] * nc[1] * nc[0], o).reshape(
                nscalars, nc[2], nc[1], nc[0]).copy()
            o += a.nbytes
            blk["s"] = a
        blocks.append(blk)
    return {"par": par, "nbtotal": nbtotal, "root_level": root_level, "region": rs,
            "time": time, "dt": dt, "ncycle": ncycle, "nghost": ng, "mhd": m,
            "nx": nx, "blocks": blocks}


def run_reference(cfg, pgen, athinput_path, overrides=None, rst_every_cycle=False,
                  keep_dir=None, timeout=3600, threads=None, hst_every_cycle=False,
                  exe=None, env_extra=None, blocks=None, rst_dcycle=None):
    """Run the reference; returns dict(dts, times, zcps, zcps_omp, rst=[paths], dir, stdout).
    exe / env_extra: run another build of the same program instead (the reference compiled
    with this repository's shim, tools/build_shim.py); blocks: parameter blocks instead of
    an athinput file."""
    exe = exe or ref_binary(cfg, pgen)
    if not os.path.isfile(exe):
        raise FileNotFoundError(exe + " (run `python oracle/build_ref.py` where "
                                "/root/reference is mounted)")
    if blocks is None:
        blocks = parse_athinput(open(athinput_path).read())
    else:
        blocks = {b: dict(kv) for b, kv in blocks.items()}
    blocks.pop("comment", None)
    apply_overrides(blocks, overrides)
    if rst_every_cycle:
        blocks["output9"] = {"file_type": "rst", "dt": "1e-300"}
    if rst_dcycle:        # a dump at cycle 0 and every rst_dcycle cycles (outputs.cpp:792)
        blocks["output9"] = {"file_type": "rst", "dcycle": str(rst_dcycle)}
    if hst_every_cycle:   # outputs/history.cpp with 17 significant digits
        blocks["output8"] = {"file_type": "hst", "dt": "1e-300", "data_format": "%24.16e"}
    if threads is not None:
        blocks["mesh"]["num_threads"] = str(threads)
    d = keep_dir or tempfile.mkdtemp(prefix="abref_")
    os.makedirs(d, exist_ok=True)
    inp = os.path.join(d, "athinput.run")
    with open(inp, "w") as f:
        f.write(dump_athinput(blocks))
    env = dict(os.environ)
    if threads is not None:
        env["OMP_NUM_THREADS"] = str(threads)
    env.update(env_extra or {})
    r = subprocess.run([exe, "-i", inp], cwd=d, capture_output=True, text=True,
                       timeout=timeout, env=env)
    if r.returncode != 0:
        raise RuntimeError("reference run failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    cyc, times, dts = [], [], []
    for m in _CYCLE_RE.finditer(r.stdout):
        cyc.append(int(m.group(1)))
        times.append(float(m.group(2)))
        dts.append(float(m.group(3)))
    out = {"cycles": cyc, "times": times, "dts": dts, "dir": d, "stdout": r.stdout}
    m = re.search(r"zone-cycles/cpu_second = ([-+0-9.eE]+)", r.stdout)
    out["zcps"] = float(m.group(1)) if m else None
    m = re.search(r"zone-cycles/omp_wsecond = ([-+0-9.eE]+)", r.stdout)
    out["zcps_omp"] = float(m.group(1)) if m else None
    m = re.search(r"omp wtime used\s*= ([-+0-9.eE]+)", r.stdout)
    out["omp_wtime"] = float(m.group(1)) if m else None
    pid = blocks["job"]["problem_id"]
    hst = os.path.join(d, pid + ".hst")
    out["hst"] = None
    if hst_every_cycle and os.path.exists(hst):
        rows = [[float(x) for x in ln.split()] for ln in open(hst) if not ln.startswith("#")]
        out["hst"] = np.array(rows)
    out["rst"] = sorted(p for p in (os.path.join(d, f) for f in os.listdir(d))
                        if re.search(re.escape(pid) + r"\.\d{5}\.rst$", p))
    return out


def cleanup(run):
    shutil.rmtree(run["dir"], ignore_errors=True)


if __name__ == "__main__":
    import sys
    cfg, pgen, inp = sys.argv[1:4]
    ov = dict(a.split("=", 1) for a in sys.argv[4:])
    res = run_reference(cfg, pgen, inp, ov, rst_every_cycle=True)
    print(res["stdout"][-1500:])
    print("dts", ["%.17e" % x for x in res["dts"][:5]])
    for p in res["rst"][:3]:
        r = read_rst(p)
        print(p, r["ncycle"], r["time"], r["dt"], r["nghost"], r["mhd"], r["blocks"][0]["u"].shape)
    print(res["dir"])
