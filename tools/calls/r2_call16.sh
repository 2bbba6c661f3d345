#!/bin/bash
cd /root/repo
O=gpurun_out/r2c16; mkdir -p $O
for v in loop0 loop1; do for args in "256 1 5 3 1" "256 1 5 3 0" "256 0 5 3 0" "128 1 10 3 1"; do echo "== $v $args"; timeout 120 scratch/fb/$v $args | grep -E "_o3|TOTAL"; done; done 2>&1 | tee $O/ppm_loop.log
