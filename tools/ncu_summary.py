#!/usr/bin/env python
"""Summarise `ncu --set full` captures of the reconstruct+Riemann kernel into
profiles/flux_ncu.json, the file bench.py reads its ncu-backed roofline keys from
(`traffic`, `fp64_pipe_util`, `dram_fraction`, `local_mem_bytes`).

  ncu -i capture.ncu-rep --page raw --csv > capture.raw.csv
  python tools/ncu_summary.py --key c5:512x512x512 [--match 'k_flux<[012], 2,'] capture.raw.csv
         [--key c4:512x512x512 other.raw.csv ...] [-o profiles/flux_ncu.json]

Every `--key` starts a new group; the launches of the CSVs that follow whose kernel name matches
`--match` (default: any k_flux) are averaged.  The source hash of the build the capture ran on is
recorded (athena-gamma_b200/build.py: kernel_hash over the device sources and flags, source_hash
over everything), so a stale summary is visible on the bench line (`captured_on_kernel_hash` vs
`kernel_hash_now`).  Existing keys of the output file are kept."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

M = {
    "duration": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "fp64": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "occ": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "lld": "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "lst": "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l2hit": "lts__t_sector_hit_rate.pct",
    "l1hit": "l1tex__t_sector_hit_rate.pct",
    "inst": "smsp__inst_executed.sum",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3,
         "msecond": 1.0, "second": 1e3}


def rows_of(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[start], rows[start + 1]
    for r in rows[start + 2:]:
        if len(r) != len(hdr):
            continue
        d = {}
        for h, u, v in zip(hdr, units, r):
            if h in M.values():
                try:
                    d[h] = float(v.replace(",", ""))*SCALE.get(u, 1.0)
                except ValueError:
                    d[h] = None
            elif h == "Kernel Name":
                d[h] = v
        yield d


def summarise(launches):
    def avg(k):
        v = [x[M[k]] for x in launches if x.get(M[k]) is not None]
        return sum(v)/len(v) if v else None
    lld, lst = avg("lld"), avg("lst")
    rd, wr = avg("dram_rd"), avg("dram_wr")
    return {"launches": len(launches),
            "kernels": sorted({x["Kernel Name"].split("(")[0] for x in launches}),
            "duration_ms_under_ncu": avg("duration"),
            "dram_bytes_per_launch": (rd + wr) if rd is not None and wr is not None else None,
            "dram_read_bytes": rd, "dram_write_bytes": wr,
            "fp64_pipe_util": (avg("fp64") or 0)/100.0 if avg("fp64") is not None else None,
            "dram_fraction": (avg("dram_pct") or 0)/100.0 if avg("dram_pct") is not None else None,
            "issue_active": (avg("issue") or 0)/100.0 if avg("issue") is not None else None,
            "achieved_occupancy": (avg("occ") or 0)/100.0 if avg("occ") is not None else None,
            "registers": avg("regs"),
            "local_mem_bytes_per_launch": 32.0*((lld or 0) + (lst or 0))
            if lld is not None or lst is not None else None,
            "l1_hit_rate": avg("l1hit"), "l2_hit_rate": avg("l2hit"),
            "warp_instructions": avg("inst")}


def main():
    import athena_gamma_b200 as ab
    args = sys.argv[1:]
    out = os.path.join(ROOT, "profiles", "flux_ncu.json")
    groups, key, match = [], None, r"k_flux"
    i = 0
    while i < len(args):
        a = args[i]
        if a == "-o":
            out = args[i + 1]; i += 2
        elif a == "--key":
            key = args[i + 1]; groups.append([key, match, []]); i += 2
        elif a == "--match":
            match = args[i + 1]
            if groups and not groups[-1][2]:
                groups[-1][1] = match
            i += 2
        else:
            groups[-1][2].append(a); i += 1
    doc = {"kernels": {}}
    if os.path.exists(out):
        doc = json.load(open(out))
    doc["srchash"] = ab.build.source_hash()
    doc["kernel_hash"] = ab.build.kernel_hash()
    for key, match, files in groups:
        rx = re.compile(match)
        launches = [r for f in files for r in rows_of(f) if rx.search(r.get("Kernel Name", ""))]
        if not launches:
            print("no launch matches %r in %s" % (match, files))
            continue
        s = summarise(launches)
        s["from"] = [os.path.relpath(os.path.abspath(f), ROOT) for f in files]
        s["match"] = match
        s["srchash"] = doc["srchash"]
        s["kernel_hash"] = doc["kernel_hash"]
        doc["kernels"][key] = s
        print(key, json.dumps(s))
    json.dump(doc, open(out, "w"), indent=1, sort_keys=True)
    print("wrote", out)


if __name__ == "__main__":
    main()
