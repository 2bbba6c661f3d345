#!/bin/bash
# 8 GPUs: default (direct exchange) vs AB_P2P=0 (NCCL send/recv), device-resident legs only
cd /root/repo
N=${1:-8}
O=gpurun_out/r2mg${N}b; mkdir -p $O
for p in 1 0; do
AB_P2P=$p AB_P2P_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$p bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu > $O/bench_${N}gpu_p2p$p.json 2> $O/bench_${N}gpu_p2p$p.err
grep "direct ghost" $O/bench_${N}gpu_p2p$p.err | sort | uniq -c | head -3
python - $O/bench_${N}gpu_p2p$p.json $p <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("p2p=" + sys.argv[2], "ngpu", d["n_gpus"], "%.4g zc/s" % d["value"], "%.3f ms" % d["ms_per_step"], "parity", d.get("parity"))
except Exception as ex:
    print("FAILED", ex)
PY
done 2>&1 | tee $O/summary.log
