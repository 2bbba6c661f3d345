#!/bin/bash
# one whole c5 cycle (23 launches) under ncu --set full at 512^3; smoke()
cd /root/repo
O=gpurun_out/r2c18; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log
B="--no-cpu --no-e2e --no-side"
# launches before the timed cycle: initialize + 1 warm-up cycle; find the skip count from the launch list
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/ll.csv python bench.py --steps 1 --warmup 1 $B > $O/ll.out 2>&1
python - $O/ll.csv <<'PY' | tee $O/skip.txt
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
names = [r[4] for r in rows]
# the timed cycle is the LAST 23 launches before the trailing history / state kernels: find the last k_flux<0, 1 launch
idx = [i for i, n in enumerate(names) if "k_flux<0, 1" in n]
print(idx[-1], len(names))
PY
S=$(cut -d' ' -f1 $O/skip.txt)
timeout 2400 ncu --set full --clock-control none -s $S -c 23 -f -o $O/prof_cycle python bench.py --steps 1 --warmup 1 $B > $O/ncu_cycle.log 2>&1
ncu -i $O/prof_cycle.ncu-rep --page raw --csv > $O/prof_cycle.raw.csv 2>/dev/null
python tools/ncu_cycle.py $O/prof_cycle.raw.csv 134217728 | tee $O/ncu_full_cycle.csv | tail -30
rm -f $O/prof_cycle.ncu-rep
