mkdir -p gpurun_out
export WANDB_MODE=disabled
L=gpurun_out/merge_exp.log
: > $L
COMMON="--steps 20 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
for env in "" "B2H_NO_TAP_MERGE=1" "B2H_NO_PERSIST=1" "B2H_NO_PERSIST=1 B2H_NO_TAP_MERGE=1"; do
  for cfg in "--mode infer --batch 4096 --frames 64" ""; do
    env $env timeout 120 python bench.py $cfg $COMMON > gpurun_out/merge_last.out 2>/dev/null
    echo "[$env] [$cfg] rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/merge_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4), 'value', round(d['value']))
except Exception as e: print('none')
")" | tee -a $L
  done
done
# per-kernel times of the persistent eval forward
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_tc -s 27 -c 9 --csv --log-file gpurun_out/persist_launches.csv \
  python bench.py --mode infer --batch 4096 --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs > /dev/null 2>&1
echo "ncu rc=$?"
