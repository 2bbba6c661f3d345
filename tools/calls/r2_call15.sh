#!/bin/bash
cd /root/repo
O=gpurun_out/r2c15; mkdir -p $O
for v in tune_x28 x1flat; do for args in "256 1 5 3 1" "256 1 5 3 0" "128 1 10 3 1" "64 1 20 3 1"; do echo "== $v $args"; timeout 120 scratch/fb/$v $args | grep -E "x1_o3|TOTAL"; done; done 2>&1 | tee $O/x1flat.log
B="--no-cpu --no-e2e --no-side"
run() { n=$1; shift; python bench.py $B "$@" > $O/bench_$n.json 2> $O/bench_$n.err; tail -2 $O/bench_$n.err; }
run c4 --workload c4 --steps 6 --warmup 3
run c4_64blk --workload c4 --steps 6 --warmup 3 --per-gpu 512,512,512 --block 128,128,128
run c3_16blk --workload c3 --steps 40 --warmup 10 --per-gpu 2048,2048,1 --block 512,512,1
run c5_256_64blk --steps 20 --warmup 5 --per-gpu 256,256,256 --block 64,64,64
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
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_fullsize.py -x -q -p no:cacheprovider 2>&1 | tail -3
