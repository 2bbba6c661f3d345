#!/bin/bash
# copy_boxes with chunk table + fastdiv: launch list at 64 blocks, bench lines
cd /root/repo
O=gpurun_out/r2c14; mkdir -p $O
B="--no-cpu --no-e2e --no-side"
run() { n=$1; shift; python bench.py $B "$@" > $O/bench_$n.json 2> $O/bench_$n.err; tail -2 $O/bench_$n.err; }
run c5 --steps 10 --warmup 3
run c5_64blk --steps 10 --warmup 3 --per-gpu 512,512,512 --block 128,128,128
run c5_256_64blk --steps 20 --warmup 5 --per-gpu 256,256,256 --block 64,64,64
run c3_16blk --workload c3 --steps 40 --warmup 10 --per-gpu 2048,2048,1 --block 512,512,1
run c4_64blk --workload c4 --steps 6 --warmup 3 --per-gpu 512,512,512 --block 128,128,128
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "%.4g zc/s" % d["value"], "%.3f ms" % d["ms_per_step"], d["gpu_launches"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done 2>&1 | tee $O/bench_summary.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c5_256_64blk.csv python bench.py --steps 2 --warmup 1 $B --per-gpu 256,256,256 --block 64,64,64 > $O/launches.out 2>&1
python tools/launchsum.py $O/launches_c5_256_64blk.csv | head -24
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_smr.py tests/test_gpu_sched.py -x -q -p no:cacheprovider 2>&1 | tail -3
