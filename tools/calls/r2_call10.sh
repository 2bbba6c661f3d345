#!/bin/bash
cd /root/repo
O=gpurun_out/r2c10; mkdir -p $O
python tools/pcie_probe.py 2>&1 | tee $O/pcie_probe.log
for v in batch2 ppm_r16 ppm_r8m3 ppm_r7; do for args in "256 1 5 3 1" "256 1 5 3 0"; do echo "== $v $args"; timeout 120 scratch/fb/$v $args | grep -E "_o3|TOTAL"; done; done 2>&1 | tee $O/ppm_variants2.log
