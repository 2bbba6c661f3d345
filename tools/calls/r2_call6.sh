#!/bin/bash
# Round 2, GPU call 6: PPM sweeps with shared reconstruction (k_flux_ppm_*) against the
# per-interface kernel in the stand-alone harness: checksums must be identical, times per launch.
cd /root/repo
O=gpurun_out/r2c6; mkdir -p $O
for v in ppm0 ppm1 ppm1_r5 ppm1_r3 ppm1_r11; do for args in "256 1 5 3 1" "256 1 5 3 0" "256 0 5 3 0"; do echo "== $v $args"; timeout 120 scratch/fb/$v $args | grep -E "_o3|TOTAL"; done; done 2>&1 | tee $O/ppm_variants.log
