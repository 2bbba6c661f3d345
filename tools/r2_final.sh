#!/bin/bash
# final evidence of the round-2 build: GPU suite, default bench line, ncu of the sweeps, one whole cycle
cd /root/repo
O=gpurun_out/r2final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=5 2>&1 | tail -12 | tee $O/gpu_suite.log
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; tail -4 $O/bench_default.err
B="--no-cpu --no-e2e --no-side"
cap() {  # name, kernel regex, skip, count, bench args...
  n=$1; k=$2; s=$3; c=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o $O/prof_$n python bench.py --steps 1 --warmup 1 $B "$@" > $O/ncu_$n.log 2>&1
  ncu -i $O/prof_$n.ncu-rep --page raw --csv > $O/prof_$n.raw.csv 2>/dev/null
  rm -f $O/prof_$n.ncu-rep
}
cap c5 k_flux 3 3
cap c4 k_flux_ppm 6 3 --workload c4
cap c3 k_flux_ppm 2 2 --workload c3
python tools/ncu_summary.py -o $O/flux_ncu.json --key c5:512x512x512 $O/prof_c5.raw.csv --key c4:512x512x512 $O/prof_c4.raw.csv --key c3:2048x2048x1 $O/prof_c3.raw.csv > $O/ncu_summary.log 2>&1; tail -3 $O/ncu_summary.log | cut -c1-300
timeout 900 ncu --set full --clock-control none -s 29 -c 23 -f -o $O/prof_cycle python bench.py --steps 1 --warmup 1 $B > $O/ncu_cycle.log 2>&1
ncu -i $O/prof_cycle.ncu-rep --page raw --csv > $O/prof_cycle.raw.csv 2>/dev/null
python tools/ncu_cycle.py $O/prof_cycle.raw.csv 134217728 | tee $O/ncu_full_cycle.csv | tail -4
rm -f $O/prof_cycle.ncu-rep
