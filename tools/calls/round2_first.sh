#!/bin/bash
# First GPU call of round 2: everything that was written at the end of round 1 without a GPU.
#   gpurun --timeout 900 -- 'bash tools/calls/round2_first.sh'
# 1. the regular GPU suite (must stay green: the one-level path was not touched)
# 2. the two xfail-marked tests with their output (device SMR path, ab_stage_* pipeline)
# 3. the isothermal-Roe fixtures without their xfail marker
# 4. bench with the pipelined e2e leg (its first_gpu_runs key repeats 2 and 3)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_gpu_suite.log
python tests/smr_check.py $(python - <<'PY'
import sys
sys.path[:0] = ["tests", "oracle", "."]
import test_gpu_smr
print(" ".join(test_gpu_smr.device_smr_goldens()))
PY
) 2>&1 | tail -20 | tee gpurun_out/r2_smr.log
python tests/stage_check.py 2>&1 | tail -5 | tee gpurun_out/r2_stage.log
# isothermal Roe (device functions written after the round-1 GPU budget was spent; verified on
# the CPU through the emulated device path): its goldens + task cases without the xfail marker
python -m pytest tests -m gpu -q --runxfail -p no:cacheprovider --tb=line -k "iso and roe" 2>&1 | tail -8 | tee gpurun_out/r2_iso_roe.log
python bench.py --steps 6 --warmup 3 --e2e-pipelined --no-cpu > gpurun_out/r2_bench_pipelined.json 2> gpurun_out/r2_bench_pipelined.err
tail -c 1500 gpurun_out/r2_bench_pipelined.json
