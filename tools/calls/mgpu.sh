#!/bin/bash
cd /root/repo
N=${1:-2}
timeout 400 python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -3
for ov in 0 1; do
AB_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-e2e 2>/dev/null | grep '^{' > gpurun_out/bench_r1_v8_${N}gpu_overlap$ov.json
python -c "
import json
d=json.loads(open('gpurun_out/bench_r1_v8_${N}gpu_overlap$ov.json').read().strip().splitlines()[-1]); print('overlap=$ov', 'ngpu', d['n_gpus'], '%.4g zc/s'%d['value'], '%.2f ms'%d['ms_per_step'])"
done
