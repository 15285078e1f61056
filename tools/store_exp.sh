mkdir -p gpurun_out
export WANDB_MODE=disabled
L=gpurun_out/store_exp.log
: > $L
COMMON="--steps 20 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
for dbg in 0 1; do
  for cfg in "" "--mode infer --batch 256 --frames 64"; do
  B2H_DBG_SKIP_STORES=$dbg timeout 120 python bench.py $cfg $COMMON > gpurun_out/store_last.out 2>/dev/null
  echo "skip_stores=$dbg [$cfg] rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/store_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4))
except Exception as e: print('none')
")" | tee -a $L
  done
  B2H_DBG_SKIP_STORES=$dbg QUIET=1 timeout 200 python tools/microbench.py "" > gpurun_out/store_micro_$dbg.txt 2>&1
  tail -13 gpurun_out/store_micro_$dbg.txt | tee -a $L
done
