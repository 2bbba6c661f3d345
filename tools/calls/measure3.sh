#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 200 python tools/stress.py c5_blast_hlld_plm_vl2_8blk,khs3d_mhd_hlld_plm_vl2_8blk_s1,c4_kh_hllc_ppm_rk2_8blk 40 2>&1 | tail -2
for ns in 0 2 4 8; do
for cfg in "256,256,256" "128,128,128"; do
AB_BLOCK_STREAMS=$ns python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e --block $cfg --per-gpu 512,512,512 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('streams $ns block $cfg', '%.4g zc/s'%d['value'], '%.2f ms'%d['ms_per_step'])
"
done; done
