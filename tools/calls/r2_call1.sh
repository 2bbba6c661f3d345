#!/bin/bash
# Round 2, GPU call 1: baseline of every BASELINE.json workload on the unchanged round-1 build
# (bench lines by CUDA events, no profiler), ncu launch lists for c3 / c4, one ncu --set full
# capture of their top kernels, compute-sanitizer memcheck over a few goldens.
cd /root/repo
O=gpurun_out/r2c1
mkdir -p $O
B="--no-cpu --no-e2e --no-first-runs"
python bench.py --steps 10 --warmup 3 $B > $O/bench_c5.json 2> $O/bench_c5.err
python bench.py --workload c2 --steps 200 --warmup 20 $B > $O/bench_c2.json 2> $O/bench_c2.err
python bench.py --workload c3 --steps 40 --warmup 10 $B > $O/bench_c3_1blk.json 2> $O/bench_c3.err
python bench.py --workload c3 --steps 40 --warmup 10 $B --block 512,512,1 > $O/bench_c3_16blk.json 2>> $O/bench_c3.err
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
done
# launch lists (shares only; cold-cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 1 $B > $O/ncu_c4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 1 $B > $O/ncu_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 1 $B > $O/ncu_c2.log 2>&1
# full capture: PPM flux kernels of c4 (hydro HLLC) and c3 (MHD HLLD), second cycle
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_flux' -s 6 -c 6 -o $O/prof_c4 -f python bench.py --workload c4 --steps 1 --warmup 1 $B > $O/ncu_full_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_flux' -s 4 -c 4 -o $O/prof_c3 -f python bench.py --workload c3 --steps 1 --warmup 1 $B > $O/ncu_full_c3.log 2>&1
# compute-sanitizer: memcheck over the c5 / c3 / smr goldens
timeout 900 compute-sanitizer --tool memcheck --target-processes all --log-file $O/memcheck.log python -m pytest tests/test_gpu_golden.py tests/test_gpu_smr.py -m gpu -q -x -p no:cacheprovider -k "c5 or c3 or smr" > $O/memcheck_pytest.log 2>&1
tail -3 $O/memcheck_pytest.log; tail -5 $O/memcheck.log
ls -la $O
