#!/bin/bash
# compute-sanitizer over the final build (after the copy-kernel / index-decode changes)
cd /root/repo
O=gpurun_out/r2finalsan; mkdir -p $O
CS="compute-sanitizer --error-exitcode 9 --print-limit 20"
K="c5_blast or c3_ot or c4_kh or c1_sod"
AB_DEBUG_ALLOC=1 timeout 400 $CS --tool memcheck python -m pytest tests/test_gpu_golden.py -x -q -p no:cacheprovider -k "$K" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc $?" | tee -a $O/sanitizer_memcheck.log
timeout 400 $CS --tool memcheck python -m pytest tests/test_gpu_golden.py -x -q -p no:cacheprovider -k "c5_blast or c3_ot or c4_kh" > $O/sanitizer_memcheck_batched.log 2>&1; echo "memcheck batched rc $?" | tee -a $O/sanitizer_memcheck_batched.log
timeout 400 $CS --tool racecheck python -m pytest tests/test_gpu_golden.py -x -q -p no:cacheprovider -k "c5_blast or c3_ot or c4_kh" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc $?" | tee -a $O/sanitizer_racecheck.log
timeout 300 $CS --tool synccheck python -m pytest tests/test_gpu_golden.py -x -q -p no:cacheprovider -k "c5_blast or c4_kh" > $O/sanitizer_synccheck.log 2>&1; echo "synccheck rc $?" | tee -a $O/sanitizer_synccheck.log
for f in $O/sanitizer_*.log; do echo "== $f"; tail -3 $f; done
