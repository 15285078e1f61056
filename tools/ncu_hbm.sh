# ncu --set full of the bandwidth-bound kernels of the bf16 training step (L1 loss, prep, Adam, bn_apply, bn_bwd), after the
# same command has exited 0 without ncu.  Report -> gpurun_out/ncu_r02_hbm.ncu-rep + raw CSV page.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export WANDB_MODE=disabled
A="bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
timeout 120 python $A > /dev/null 2>&1 && \
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"l1_vec_kernel|prep_ncl_vec_kernel|adam_kernel|bn_apply_kernel|bn_bwd_kernel" \
  -s 120 -c 24 -o gpurun_out/ncu_r02_hbm -f python $A > gpurun_out/ncu_r02_hbm.log 2>&1
echo "hbm rc=$?"
ncu -i gpurun_out/ncu_r02_hbm.ncu-rep --page raw --csv > gpurun_out/ncu_r02_hbm_raw.csv 2>/dev/null
ls -la gpurun_out | grep ncu_r02_hbm
