#!/usr/bin/env python3
"""Golden L1 errors of the linear-wave convergence test, produced by the UNMODIFIED reference
(oracle/_ref/mhd_hlld_ng2/athena_linear_wave, problem/compute_error=true ->
linearwave-errors.dat, src/pgen/linear_wave.cpp:190-428) at two resolutions: BASELINE
configs[1] (128x64x64) and half of it.  north_star: the product must reproduce these to three
significant digits.  Writes tests/golden/linwave_errors.json."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_run  # noqa: E402

out = {}
for res, blk in (((64, 32, 32), (32, 16, 16)), ((128, 64, 64), (64, 32, 32))):
    ov = {"mesh/nx1": res[0], "mesh/nx2": res[1], "mesh/nx3": res[2],
          "meshblock/nx1": blk[0], "meshblock/nx2": blk[1], "meshblock/nx3": blk[2],
          "time/ncycle_out": 0}
    r = ref_run.run_reference("mhd_hlld_ng2", "linear_wave",
                              os.path.join(ROOT, "inputs", "athinput.linear_wave3d"), ov,
                              threads=8)
    rows = [ln.split() for ln in open(os.path.join(r["dir"], "linearwave-errors.dat"))
            if not ln.startswith("#")]
    v = rows[-1]
    out["x".join(map(str, res))] = {
        "ncycle": int(v[3]), "rms": float(v[4]), "l1": [float(x) for x in v[5:13]],
        "overrides": ov}
    ref_run.cleanup(r)
    print(res, out["x".join(map(str, res))]["ncycle"], out["x".join(map(str, res))]["rms"])
json.dump(out, open(os.path.join(HERE, "linwave_errors.json"), "w"), indent=1)
