mkdir -p gpurun_out
export WANDB_MODE=disabled
L=gpurun_out/e2e_var.log
: > $L
cat /sys/bus/pci/devices/*/local_cpulist 2>/dev/null | sort | uniq -c | head -5 | tee -a $L
nproc | tee -a $L
COMMON="--steps 20 --warmup 5 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
for i in 1 2 3 4 5 6; do
  timeout 120 python bench.py $COMMON > gpurun_out/e2e_last.out 2>/dev/null
  echo "run $i rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/e2e_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['e2e']['runs_ms'], d['e2e']['host_affinity'])
except Exception as e: print('none', e)
")" | tee -a $L
done
