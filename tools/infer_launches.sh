mkdir -p gpurun_out
export WANDB_MODE=disabled
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:bn_apply|bn_fold|gemm_tc|pack_multi|prep_ncl|to_ncl" -s 50 -c 16 --csv --log-file gpurun_out/infer_launches.csv \
  python bench.py --mode infer --batch 4096 --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs > /dev/null 2>&1
echo "ncu rc=$?"
python - <<PY
import csv
lines=[l for l in open('gpurun_out/infer_launches.csv') if not l.startswith('==')]
for r in csv.DictReader(lines):
    print(round(float(r['Metric Value'])/1000,1), r['Kernel Name'][:70], r['Grid Size'])
PY
