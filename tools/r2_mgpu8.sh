#!/bin/bash
# Round 2, 8-GPU validation: the default bench line under torchrun (every leg incl. e2e, parity)
cd /root/repo
N=${1:-8}
O=gpurun_out/r2mg$N; mkdir -p $O
nvidia-smi -L | wc -l; free -g | head -2
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu ) > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
tail -5 $O/bench_${N}gpu.err
python - $O/bench_${N}gpu.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("ngpu", d["n_gpus"], "%.4g zc/s" % d["value"], "%.2f ms" % d["ms_per_step"], "parity", d.get("parity"),
          "e2e", (d.get("e2e") or {}).get("value"), "plain", (d.get("e2e_plain") or {}))
except Exception as ex:
    print("FAILED", ex)
PY
