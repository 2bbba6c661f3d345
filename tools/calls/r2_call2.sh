#!/bin/bash
# Round 2, GPU call 2: flux-kernel variants in the stand-alone harness (tools/fluxbench.cu);
# every variant must print the checksum of `base` (the round-1 kernel) for the same arguments.
cd /root/repo
O=gpurun_out/r2c2; mkdir -p $O
for v in base v96 v128 v80 v64x9 v112 v104; do
  for args in "256 0 5 2 0" "256 1 5 2 0"; do
    echo "== $v $args"; timeout 120 scratch/fb/$v $args | tail -1
  done
done 2>&1 | tee $O/variants.log
for v in base v96; do
  for args in "512 0 3 2 0" "256 0 5 3 1" "256 1 5 3 1" "256 1 5 3 0" "256 0 5 3 0"; do
    echo "== $v $args"; timeout 120 scratch/fb/$v $args
  done
done 2>&1 | tee $O/base_vs_v96.log
