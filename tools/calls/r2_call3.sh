#!/bin/bash
# Round 2, GPU call 3: the rebuilt library (mirrored HLLD, lazy PPM extremum fix, multiply-high
# index division) through the whole GPU suite, bench lines for every BASELINE workload, and an
# ncu --set full capture of the HLLD sweep in the stand-alone harness.
cd /root/repo
O=gpurun_out/r2c3; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -6 | tee $O/gpu_suite.log
B="--no-cpu --no-e2e"
python bench.py --steps 10 --warmup 3 $B > $O/bench_c5.json 2> $O/bench_c5.err
python bench.py --workload c2 --steps 200 --warmup 20 $B > $O/bench_c2.json 2> $O/bench_c2.err
python bench.py --workload c3 --steps 40 --warmup 10 $B > $O/bench_c3_1blk.json 2> $O/bench_c3.err
python bench.py --workload c3 --steps 40 --warmup 10 $B --per-gpu 2048,2048,1 --block 512,512,1 > $O/bench_c3_16blk.json 2>> $O/bench_c3.err
python bench.py --workload c4 --steps 6 --warmup 3 $B > $O/bench_c4.json 2> $O/bench_c4.err
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(sys.argv[1], "%.4g zc/s" % d["value"], "%.3f ms" % d["ms_per_step"], d["gpu_launches"],
          {k: round(v, 3) for k, v in (r.get("flux_avg_ms_by_dir_order") or {}).items()})
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done 2>&1 | tee $O/bench_summary.log
for sel in "0 x1_o1" "9 x1_o2"; do set -- $sel
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_flux -s $1 -c 1 -f -o $O/prof_fb_$2 scratch/fb/v96 256 0 1 2 0 > $O/ncu_fb_$2.log 2>&1
  ncu -i $O/prof_fb_$2.ncu-rep --page raw --csv > $O/prof_fb_$2.raw.csv 2>/dev/null
  ncu -i $O/prof_fb_$2.ncu-rep --page source --csv > $O/prof_fb_$2.source.csv 2>/dev/null
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_flux -s 9 -c 1 -f -o $O/prof_fb_x1_o2_dev scratch/fb/v96 256 1 1 2 0 > $O/ncu_fb_dev.log 2>&1
ncu -i $O/prof_fb_x1_o2_dev.ncu-rep --page raw --csv > $O/prof_fb_x1_o2_dev.raw.csv 2>/dev/null
ls -la $O; du -sh $O
