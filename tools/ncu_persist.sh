# ncu --set full of the persistent inference GEMMs (one forward at 4096 x 64 bf16) + smoke()
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export WANDB_MODE=disabled
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
A="bench.py --mode infer --batch 4096 --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
timeout 120 python $A > /dev/null 2>&1 && \
timeout 500 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_persist -s 27 -c 9 -o gpurun_out/ncu_r02_persist -f \
  python $A > gpurun_out/ncu_r02_persist.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/ncu_r02_persist.ncu-rep --page raw --csv > gpurun_out/ncu_r02_persist_raw.csv 2>/dev/null
ls -la gpurun_out | grep ncu_r02_persist
