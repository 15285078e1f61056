# persistent multi-tile GEMM (k_gemm_persist.cu): replay parity at 64 x 1024, eval parity at >= 32768 frames (ragged),
# then the inference shapes with / without it
mkdir -p gpurun_out
export WANDB_MODE=disabled
L=gpurun_out/persist.log
: > $L
timeout 300 python -m pytest tests/test_gpu_replay.py -x -q -p no:cacheprovider -k "64-1024 or test_discriminator_eval_replay" >> $L 2>&1
echo "tests rc=$?" | tee -a $L
grep -E "passed|failed" $L | tail -2
timeout 400 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "test_eval_forward_vs_oracle" >> $L 2>&1
echo "parity rc=$?" | tee -a $L
grep -E "passed|failed" $L | tail -1
COMMON="--steps 20 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
for cfg in "--mode infer --batch 4096 --frames 64" "--mode infer --batch 64 --frames 1024" "--mode infer --batch 1024 --frames 64" \
           "--mode infer --variant v2 --feats --batch 4096 --frames 64" ""; do
  for np in 0 1; do
    if [ "$np" = "1" ]; then export B2H_NO_PERSIST=1 B2H_NO_NCL_DIRECT=1; else unset B2H_NO_PERSIST B2H_NO_NCL_DIRECT; fi
    timeout 120 python bench.py $cfg $COMMON > gpurun_out/persist_last.out 2>/dev/null
    echo "no_persist=$np [$cfg] rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/persist_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d.get('e2e',{}).get('value',0)))
except Exception as e: print('none')
")" | tee -a $L
  done
done
