mkdir -p gpurun_out
export WANDB_MODE=disabled
L=gpurun_out/epi_exp.log
: > $L
COMMON="--steps 20 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
for dbg in 0 1 2; do
  B2H_PERSIST_DBG=$dbg timeout 120 python bench.py --mode infer --batch 4096 --frames 64 $COMMON > gpurun_out/epi_last.out 2>/dev/null
  echo "dbg=$dbg rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/epi_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4))
except Exception as e: print('none')
")" | tee -a $L
  B2H_PERSIST_DBG=$dbg timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_tc -s 27 -c 9 --csv --log-file gpurun_out/epi_launches_$dbg.csv \
    python bench.py --mode infer --batch 4096 --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs > /dev/null 2>&1
  python - <<PY | tee -a $L
import csv
lines=[l for l in open('gpurun_out/epi_launches_$dbg.csv') if not l.startswith('==')]
print('  per-kernel us:', [round(float(r['Metric Value'])/1000,1) for r in csv.DictReader(lines)])
PY
done
