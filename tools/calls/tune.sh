#!/bin/bash
cd /root/repo
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$name', 'ms/step %.2f'%d['ms_per_step'], 'flux', {k:round(v,2) for k,v in r['flux_avg_ms_by_dir_order'].items()})
"
}
run default AB_X=1
for f in tools/libs/lib_*.so; do run $(basename $f .so) AB_LIB=$PWD/$f; done
