#!/bin/bash
# integrate_fc tile mapping: launch list (kernel time) and cycle time for the tile variants
cd /root/repo
O=gpurun_out/r2c13; mkdir -p $O
B="--no-cpu --no-e2e --no-side"
for v in default fc_8x2 fc_4x4 fc_2x2; do
  if [ $v = default ]; then unset AB_LIB; else export AB_LIB=/root/repo/scratch/libs/$v.so; fi
  python bench.py $B --steps 10 --warmup 3 > $O/bench_$v.json 2> $O/bench_$v.err
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_$v.csv python bench.py --steps 2 --warmup 1 $B > $O/launches_$v.out 2>&1
  echo "== $v"; python -c "
import json; d=json.loads(open('$O/bench_$v.json').read().strip().splitlines()[-1]); print('%.4g zc/s %.3f ms'%(d['value'], d['ms_per_step']))"
  python tools/launchsum.py $O/launches_$v.csv | grep -E "integrate_fc|integrate_cc|corner"
done 2>&1 | tee $O/summary.log
