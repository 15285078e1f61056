# tile-width sweep (B2H_FORCE_BN) of the eval forward at the inference shapes and of the training step
mkdir -p gpurun_out
export WANDB_MODE=disabled
L=gpurun_out/bn_sweep.log
: > $L
COMMON="--steps 20 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
for bn in 0 256 128 64; do
  for cfg in "--mode infer --batch 4096 --frames 64" "--mode infer --batch 64 --frames 1024" "--mode infer --batch 256 --frames 64" ""; do
    if [ "$bn" = "0" ]; then unset B2H_FORCE_BN; else export B2H_FORCE_BN=$bn; fi
    timeout 100 python bench.py $cfg $COMMON > gpurun_out/bn_sweep_last.out 2>/dev/null
    echo "BN=$bn [$cfg] rc=$? $(python -c "
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/bn_sweep_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4), 'value', round(d['value']))
except Exception as e: print('none')
")" | tee -a $L
  done
done
