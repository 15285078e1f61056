mkdir -p gpurun_out
export WANDB_MODE=disabled
L=gpurun_out/pair_quick.log
: > $L
B2H_PAIR=1 timeout 150 python -m pytest tests/test_gpu_replay.py -x -q -p no:cacheprovider -k "256-64-1 or 64-1024-1 or test_forced_tile_widths_replay" >> $L 2>&1
echo "tests rc=$?" | tee -a $L
COMMON="--steps 20 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
for cfg in "" "--mode infer --batch 4096 --frames 64" "--mode infer --batch 64 --frames 1024" "--mode infer --batch 256 --frames 64"; do
  for pair in 1 0; do
    B2H_PAIR=$pair timeout 120 python bench.py $cfg $COMMON > gpurun_out/pair_last.out 2>/dev/null
    echo "pair=$pair [$cfg] rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/pair_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4), 'value', round(d['value']))
except Exception as e: print('none')
")" | tee -a $L
  done
done
