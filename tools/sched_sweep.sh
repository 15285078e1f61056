mkdir -p gpurun_out
export WANDB_MODE=disabled
L=gpurun_out/sched_sweep.log
: > $L
COMMON="--steps 30 --warmup 5 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
for env in "" "B2H_WGRAD_STREAMS=2" "B2H_WGRAD_STREAMS=3" "B2H_WGRAD_STREAMS=6" "B2H_WGRAD_STREAMS=8" "B2H_WGRAD_DIRECT_SKIP=1" "B2H_WGRAD_DIRECT_SKIP=3" "B2H_WGRAD_DIRECT_SKIP=4" "B2H_BUCKETS=2" "B2H_BUCKETS=4" "B2H_WGRAD_MAX_WN=128" "B2H_NO_WGRAD_DIRECT=1"; do
  env $env timeout 120 python bench.py $COMMON > gpurun_out/sched_last.out 2>/dev/null
  echo "[$env] rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/sched_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['runs_ms'][1]/30,4))
except Exception as e: print('none')
")" | tee -a $L
done
