#!/bin/bash
cd /root/repo
O=gpurun_out/r2c11; mkdir -p $O
for v in tune0 tune_x24 tune_x28 tune_h4 tune_m6; do for args in "256 1 5 3 1" "256 1 5 3 0"; do echo "== $v $args"; timeout 120 scratch/fb/$v $args | grep -E "_o3|TOTAL"; done; done 2>&1 | tee $O/ppm_variants3.log
