#!/bin/bash
# Round 2, multi-GPU call: 2-rank golden parity (both schedules), then bench.py under torchrun
# with the in-order and the overlapped schedule (the line carries `parity`).
cd /root/repo
N=${1:-2}
O=gpurun_out/r2mg$N; mkdir -p $O
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -p no:cacheprovider 2>&1 | tail -4 | tee $O/multirank_tests.log
for ov in 0 1; do
AB_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_${N}gpu_overlap$ov.json 2> $O/bench_${N}gpu_overlap$ov.err
python - $O/bench_${N}gpu_overlap$ov.json $ov <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("overlap=" + sys.argv[2], "ngpu", d["n_gpus"], "%.4g zc/s" % d["value"], "%.2f ms" % d["ms_per_step"],
          "parity", d.get("parity"), "e2e", (d.get("e2e") or {}).get("value"))
except Exception as ex:
    print("FAILED", ex)
PY
tail -3 $O/bench_${N}gpu_overlap$ov.err
done 2>&1 | tee $O/summary.log
