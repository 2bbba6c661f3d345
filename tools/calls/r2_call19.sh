#!/bin/bash
cd /root/repo
O=gpurun_out/r2c19; mkdir -p $O
B="--no-cpu --no-e2e --no-side"
python bench.py $B --steps 10 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err
python bench.py $B --workload c4 --steps 6 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py $B --steps 20 --warmup 5 --per-gpu 256,256,256 --block 64,64,64 > $O/bench_c5_256_64blk.json 2> $O/bench_c5_256.err
for n in c5 c4 c5_256_64blk; do python -c "
import json; d=json.loads(open('$O/bench_$n.json').read().strip().splitlines()[-1]); print('$n %.4g zc/s %.3f ms'%(d['value'], d['ms_per_step']))"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_c5.csv python bench.py --steps 2 --warmup 1 $B > $O/launches.out 2>&1
python tools/launchsum.py $O/launches_c5.csv | grep -E "integrate|corner|cons2prim"
