#!/bin/bash
# Round 2, GPU call 4: whole GPU suite (new: polling scheduler, reference+shim binaries, full-size
# bitwise tests vs the live reference), fdiv variant in the flux harness, the shimmed reference
# binary at 256^3 (its own zone-cycles/s printout) beside the unmodified reference.
cd /root/repo
O=gpurun_out/r2c4; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=8 2>&1 | tail -25 | tee $O/gpu_suite.log
for v in v96_prev v96; do for args in "256 0 5 2 0" "256 1 5 2 0" "256 0 5 3 1" "256 1 5 3 0"; do echo "== $v $args"; timeout 120 scratch/fb/$v $args | tail -1; done; done 2>&1 | tee $O/fdiv.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_c5.json 2> $O/bench_c5.err; tail -c 900 $O/bench_c5.json
# the drop-in at scale: 256^3 MHD blast in 8 MeshBlocks of 128^3, 20 cycles, no outputs
mkdir -p $O/shimrun && cd $O/shimrun
OV="mesh/nx1=256 mesh/nx2=256 mesh/nx3=256 meshblock/nx1=128 meshblock/nx2=128 meshblock/nx3=128 time/nlim=20 time/tlim=1e30 time/ncycle_out=5"
IN=/root/repo/inputs/athinput.blast
( time /root/repo/shim/_build/mhd_hlld_ng2/athena_blast -i $IN $OV ) > shim.log 2>&1; tail -12 shim.log
( time /root/repo/shim/_build/mhd_hlld_ng2/athena_blast -i $IN $OV mesh/num_threads=4 ) > shim_t4.log 2>&1; tail -6 shim_t4.log
( time /root/repo/oracle/_ref/mhd_hlld_ng2/athena_blast -i $IN $OV mesh/num_threads=8 time/nlim=4 ) > ref.log 2>&1; tail -6 ref.log
rm -f *.rst *.hst
cd /root/repo; du -sh $O
