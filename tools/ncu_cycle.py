#!/usr/bin/env python
"""One whole cycle under `ncu --set full`: per-launch table and the cycle's DRAM traffic.

  ncu -i cycle.ncu-rep --page raw --csv > cycle.raw.csv
  python tools/ncu_cycle.py cycle.raw.csv zones > profiles/rN_ncu_full_cycle.csv
"""
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import ncu_summary as ns  # noqa: E402


def main():
    path, zones = sys.argv[1], float(sys.argv[2])
    rows = list(ns.rows_of(path))
    M = ns.M
    print("kernel,ms,dram_rd_GB,dram_wr_GB,dram_pct,fp64_pipe_pct,issue_pct,regs,occupancy_pct,"
          "l1_hit_pct,l2_hit_pct,local_mem_GB")
    tot_ms = tot_rd = tot_wr = 0.0
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        ms, rd, wr = r[M["duration"]], r[M["dram_rd"]], r[M["dram_wr"]]
        loc = 32.0*((r.get(M["lld"]) or 0) + (r.get(M["lst"]) or 0))
        tot_ms += ms; tot_rd += rd; tot_wr += wr
        print('"%s",%.4f,%.3f,%.3f,%.1f,%.1f,%.1f,%d,%.1f,%.1f,%.1f,%.2f' %
              (name, ms, rd/1e9, wr/1e9, r[M["dram_pct"]], r[M["fp64"]], r[M["issue"]],
               r[M["regs"]], r[M["occ"]], r[M["l1hit"]], r[M["l2hit"]], loc/1e9))
    print("# %d launches, %.2f ms under ncu (cold cache, serialised), DRAM read %.1f GB + write "
          "%.1f GB = %.0f bytes per zone-cycle (%.3g zones)" %
          (len(rows), tot_ms, tot_rd/1e9, tot_wr/1e9, (tot_rd + tot_wr)/zones, zones))


if __name__ == "__main__":
    main()
