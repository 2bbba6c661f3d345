#!/bin/bash
# round-1 measurement set: bench lines (CUDA events, no profiler) + ncu launch list + ncu full
cd /root/repo
O=gpurun_out
python bench.py --steps 10 --warmup 3 > $O/bench_r1_v9.json 2> $O/bench_r1_v9.err
tail -c 600 $O/bench_r1_v9.json
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --block 128,128,128 --per-gpu 512,512,512 > $O/bench_r1_v9_64blk.json 2>> $O/bench_r1_v9.err
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --block 256,256,256 --per-gpu 512,512,512 > $O/bench_r1_v9_8blk.json 2>> $O/bench_r1_v9.err
for f in $O/bench_r1_v9_64blk.json $O/bench_r1_v9_8blk.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['gpu_launches'])"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1_v9_512cube.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $O/b_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_flux|k_integrate_cc|k_integrate_fc|k_corner_e3d|k_cons2prim' -s 17 -c 14 -o $O/prof_r1_v9_512cube -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $O/b_ncu_full.log 2>&1
tail -3 $O/b_ncu_full.log
ls -la $O | tail -8
