#!/bin/bash
# last GPU visit of the round (about two minutes of budget): the deferred-tail variants on the device, an A/B of the
# default bench, then as much of the -m gpu suite as fits
mkdir -p gpurun_out
export WANDB_MODE=disabled
LOG=gpurun_out/ab_r02x6.log
: > $LOG
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "opt_in_backward" 2>&1 | tail -2 | tee -a $LOG
COMMON="--steps 20 --warmup 5 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
for v in "B2H_DEFER_BN=2" "B2H_DEFER_BN=0" "B2H_DEFER_BN=2"; do
  env $v timeout 60 python bench.py $COMMON > gpurun_out/ab_last.out 2> gpurun_out/ab_last.err
  echo "[$v] rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/ab_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e6,2))
except Exception as e: print('none', e)
")" | tee -a $LOG
done
timeout 100 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_r02x6.log 2>&1
echo "pytest rc=$?" | tee -a $LOG
tail -3 gpurun_out/pytest_gpu_r02x6.log | tee -a $LOG
