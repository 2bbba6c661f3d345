#!/bin/bash
cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in "512,512,512" "256,256,256" "128,128,128"; do
python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e --block $cfg --per-gpu 512,512,512 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('block $cfg', '%.4g zc/s'%d['value'], '%.2f ms'%d['ms_per_step'], d['gpu_launches'], {k:round(v,2) for k,v in d['roofline']['flux_avg_ms_by_dir_order'].items()})
"
done
