#!/bin/bash
# Round 2, GPU call 7: one launch per task over all MeshBlocks (ab_batch.cuh) + shared PPM
# reconstruction in the library: harness A/B of the pointer shift in k_flux, whole GPU suite,
# bench lines of many-block meshes (AB_NO_BATCH=1 = block-by-block launches as before).
cd /root/repo
O=gpurun_out/r2c7; mkdir -p $O
for v in ppm1 batch1; do for args in "256 0 5 2 0" "256 1 5 2 0" "256 1 5 3 1"; do echo "== $v $args"; timeout 120 scratch/fb/$v $args | tail -1; done; done 2>&1 | tee $O/shift_ab.log
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 | tee $O/gpu_suite.log
B="--no-cpu --no-e2e --no-side"
run() { n=$1; shift; python bench.py $B "$@" > $O/bench_$n.json 2> $O/bench_$n.err; tail -2 $O/bench_$n.err; }
run c5 --steps 10 --warmup 3
run c5_64blk --steps 10 --warmup 3 --per-gpu 512,512,512 --block 128,128,128
AB_NO_BATCH=1 run c5_64blk_nobatch --steps 10 --warmup 3 --per-gpu 512,512,512 --block 128,128,128
run c5_256_64blk --steps 20 --warmup 5 --per-gpu 256,256,256 --block 64,64,64
AB_NO_BATCH=1 run c5_256_64blk_nobatch --steps 20 --warmup 5 --per-gpu 256,256,256 --block 64,64,64
run c3_16blk --workload c3 --steps 40 --warmup 10 --per-gpu 2048,2048,1 --block 512,512,1
AB_NO_BATCH=1 run c3_16blk_nobatch --workload c3 --steps 40 --warmup 10 --per-gpu 2048,2048,1 --block 512,512,1
run c3_1blk --workload c3 --steps 40 --warmup 10
run c4 --workload c4 --steps 6 --warmup 3
run c4_64blk --workload c4 --steps 6 --warmup 3 --per-gpu 512,512,512 --block 128,128,128
run c2 --workload c2 --steps 200 --warmup 20
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(sys.argv[1], "%.4g zc/s" % d["value"], "%.3f ms" % d["ms_per_step"], d["gpu_launches"],
          {k: round(v, 3) for k, v in (r.get("flux_avg_ms_by_dir_order") or {}).items()})
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done 2>&1 | tee $O/bench_summary.log
