#!/bin/bash
# Round 2, GPU call 12 (evidence of the current build): whole GPU suite, default bench line with
# every leg (timed) and the reference arm, launch lists, ncu --set full of the Riemann sweeps of
# c5 / c4 / c3 inside bench.py, compute-sanitizer over goldens incl. the shared-memory PPM kernels.
cd /root/repo
O=gpurun_out/${1:-r2c12}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=5 2>&1 | tail -12 | tee $O/gpu_suite.log
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; tail -4 $O/bench_default.err
( time python bench.py --impl reference --steps 10 --warmup 3 ) > $O/bench_reference.json 2> $O/bench_reference.err; tail -4 $O/bench_reference.err
B="--no-cpu --no-e2e --no-side"
python bench.py $B --workload c4 --steps 6 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py $B --workload c3 --steps 40 --warmup 10 --per-gpu 2048,2048,1 --block 512,512,1 > $O/bench_c3_16blk.json 2> $O/bench_c3.err
python bench.py $B --workload c2 --steps 200 --warmup 20 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c5.csv python bench.py --steps 2 --warmup 1 $B > $O/launches_c5.out 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 1 $B > $O/launches_c4.out 2>&1
cap() {  # name, kernel regex, skip, count, bench args...
  n=$1; k=$2; s=$3; c=$4; shift 4
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o $O/prof_$n python bench.py --steps 1 --warmup 1 $B "$@" > $O/ncu_$n.log 2>&1
  ncu -i $O/prof_$n.ncu-rep --page raw --csv > $O/prof_$n.raw.csv 2>/dev/null
}
cap c5 k_flux 3 3
cap c4 k_flux_ppm 6 3 --workload c4
cap c3 k_flux_ppm 2 2 --workload c3
python tools/ncu_summary.py -o $O/flux_ncu.json --key c5:512x512x512 $O/prof_c5.raw.csv --key c4:512x512x512 $O/prof_c4.raw.csv --key c3:2048x2048x1 $O/prof_c3.raw.csv > $O/ncu_summary.log 2>&1; tail -5 $O/ncu_summary.log | cut -c1-700
ncu -i $O/prof_c4.ncu-rep --page source --csv > $O/prof_c4.source.csv 2>/dev/null
rm -f $O/prof_c4.ncu-rep $O/prof_c3.ncu-rep $O/prof_c5.ncu-rep
export AB_DEBUG_ALLOC=1
CS="compute-sanitizer --error-exitcode 9 --print-limit 20"
timeout 900 $CS --tool memcheck python -m pytest tests/test_gpu_golden.py -x -q -p no:cacheprovider -k "c5_blast or c3_ot or c4_kh or c1_sod" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc $?" | tee -a $O/sanitizer_memcheck.log
unset AB_DEBUG_ALLOC
# the batched launches need the single allocation: memcheck again without AB_DEBUG_ALLOC
timeout 900 $CS --tool memcheck python -m pytest tests/test_gpu_golden.py -x -q -p no:cacheprovider -k "c5_blast or c3_ot or c4_kh" > $O/sanitizer_memcheck_batched.log 2>&1; echo "memcheck batched rc $?" | tee -a $O/sanitizer_memcheck_batched.log
timeout 600 $CS --tool memcheck python tests/smr_check.py smr_blast2d_hllc_plm_vl2 smr_blast3d_hllc_plm_vl2 smr_khs2d_lhllc_plm_vl2_s1 > $O/sanitizer_memcheck_smr.log 2>&1; echo "memcheck smr rc $?" | tee -a $O/sanitizer_memcheck_smr.log
timeout 900 $CS --tool racecheck python -m pytest tests/test_gpu_golden.py -x -q -p no:cacheprovider -k "c5_blast or c3_ot or c4_kh" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc $?" | tee -a $O/sanitizer_racecheck.log
timeout 900 $CS --tool synccheck python -m pytest tests/test_gpu_golden.py -x -q -p no:cacheprovider -k "c5_blast or c3_ot or c4_kh" > $O/sanitizer_synccheck.log 2>&1; echo "synccheck rc $?" | tee -a $O/sanitizer_synccheck.log
timeout 600 $CS --tool initcheck python -m pytest tests/test_gpu_golden.py -x -q -p no:cacheprovider -k "c5_blast or c4_kh" > $O/sanitizer_initcheck.log 2>&1; echo "initcheck rc $?" | tee -a $O/sanitizer_initcheck.log
for f in $O/sanitizer_*.log; do echo "== $f"; tail -4 $f; done
du -sh $O
