mkdir -p gpurun_out
export WANDB_MODE=disabled
L=gpurun_out/quick.log
: > $L
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -p no:cacheprovider -m gpu -k "loss or persistent or gan_step or pipelined or golden" >> $L 2>&1
echo "tests rc=$?" | tee -a $L
grep -E "passed|failed" $L | tail -2
COMMON="--steps 20 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
for env in "" "B2H_NO_L1_FUSE=1"; do
  env $env timeout 120 python bench.py $COMMON > gpurun_out/quick_last.out 2>/dev/null
  echo "[$env] rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/quick_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4), 'e2e', round(d.get('e2e',{}).get('value',0)), d.get('e2e',{}).get('runs_ms'))
except Exception as e: print('none')
")" | tee -a $L
  env $env QUIET=1 timeout 200 python tools/microbench.py "" 2>&1 | grep -E "L1|ToNcl|ops of one" | tee -a $L
done
