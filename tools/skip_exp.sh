#!/bin/bash
# Timing-only experiment: the default bench with classes of ops left out of the step (B2H_DBG_SKIP_OPS, results invalid
# by construction) -- an upper bound of what fusing those launches away could gain.  usage: bash tools/skip_exp.sh TAG
TAG=${1:-skip}
mkdir -p gpurun_out
export WANDB_MODE=disabled
LOG=gpurun_out/skip_$TAG.log
: > $LOG
COMMON="--steps 20 --warmup 5 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
run() {
  env B2H_DBG_SKIP_OPS="$1" timeout 150 python bench.py $COMMON > gpurun_out/skip_last.out 2> gpurun_out/skip_last.err
  echo "[$1] rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/skip_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4))
except Exception as e: print('none', e)
")" | tee -a $LOG
}
run 'no-such-op'
run 'convs\.(9|13|17|21|25|29)($|\[|\.)'
run '^wgrad\.'
run '^bn_bwd\.'
run '^apply\.'
run '^(wgrad|bn_bwd|apply)\.'
