#!/bin/bash
# One short GPU-box visit: the -m gpu tests, then the default bench under a list of environment variants.
# usage (under gpurun, from the repo root): bash tools/ab_call.sh TAG "ENV1=a ENV2=b" "ENV3=c" ...   ("-" = no variables)
TAG=${1:-ab}; shift
mkdir -p gpurun_out
export WANDB_MODE=disabled
LOG=gpurun_out/ab_$TAG.log
: > $LOG
if [ -z "$AB_SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider ${AB_PYTEST_ARGS} > gpurun_out/pytest_gpu_$TAG.log 2>&1
  echo "pytest rc=$?" | tee -a $LOG
  tail -6 gpurun_out/pytest_gpu_$TAG.log | tee -a $LOG
fi
COMMON="--steps 20 --warmup 5 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
for v in "$@"; do
  [ "$v" = "-" ] && v=""
  env $v timeout 150 python bench.py $COMMON > gpurun_out/ab_last.out 2> gpurun_out/ab_last.err
  echo "[$v] rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/ab_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4), 'Mfps', round(d['value']/1e6,2), 'e2e', round(d.get('e2e',{}).get('value',0)/1e6,2))
except Exception as e: print('none', e)
")" | tee -a $LOG
done
if [ -n "$AB_MICRO" ]; then
  for v in $AB_MICRO; do
    [ "$v" = "-" ] && v=""
    echo "== microbench [$v]" | tee -a $LOG
    env $v QUIET=1 timeout 200 python tools/microbench.py "" 2>&1 | tail -14 | tee -a $LOG
  done
fi
