#!/bin/bash
# Round 2, GPU call 8: after the EMF-plan fix (plans read from the device array) and the late
# output-pointer shift in the Riemann kernels: harness A/B, bench lines again.
cd /root/repo
O=gpurun_out/r2c9; mkdir -p $O
echo skip harness
B="--no-cpu --no-e2e --no-side"
run() { n=$1; shift; python bench.py $B "$@" > $O/bench_$n.json 2> $O/bench_$n.err; tail -2 $O/bench_$n.err; }
run c5 --steps 10 --warmup 3
run c5_64blk --steps 10 --warmup 3 --per-gpu 512,512,512 --block 128,128,128
run c5_256_64blk --steps 20 --warmup 5 --per-gpu 256,256,256 --block 64,64,64
AB_NO_BATCH=1 run c5_256_64blk_nobatch --steps 20 --warmup 5 --per-gpu 256,256,256 --block 64,64,64
run c3_16blk --workload c3 --steps 40 --warmup 10 --per-gpu 2048,2048,1 --block 512,512,1
AB_NO_BATCH=1 run c3_16blk_nobatch --workload c3 --steps 40 --warmup 10 --per-gpu 2048,2048,1 --block 512,512,1
run c3_1blk --workload c3 --steps 40 --warmup 10
run c2 --workload c2 --steps 200 --warmup 20
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
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c5_256_64blk.csv python bench.py --steps 2 --warmup 1 $B --per-gpu 256,256,256 --block 64,64,64 > $O/launches.out 2>&1
python tools/launchsum.py $O/launches_c5_256_64blk.csv | head -24
