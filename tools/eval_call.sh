# eval-plan changes: replay / parity / golden / shapes on the GPU, then the inference shapes and the training step
mkdir -p gpurun_out
export WANDB_MODE=disabled
L=gpurun_out/eval_call.log
: > $L
timeout 900 python -m pytest tests/test_gpu_replay.py tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_shapes.py tests/test_fk.py tests/test_modelzoo_shim.py -q -p no:cacheprovider -m gpu >> $L 2>&1
echo "tests rc=$?" | tee -a $L
grep -E "passed|failed" $L | tail -2
COMMON="--steps 20 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
for cfg in "--mode infer --batch 4096 --frames 64" "--mode infer --batch 64 --frames 1024" "--mode infer --batch 256 --frames 64" \
           "--mode infer --variant v2 --feats --batch 4096 --frames 64" "--mode infer --batch 4096 --frames 64 --precision fp32" "" "--precision fp32"; do
  timeout 120 python bench.py $cfg $COMMON > gpurun_out/eval_last.out 2>/dev/null
  echo "[$cfg] rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/eval_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d.get('e2e',{}).get('value',0)), d.get('e2e',{}).get('runs_ms'))
except Exception as e: print('none')
")" | tee -a $L
done
