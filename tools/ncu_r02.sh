# ncu --set full captures of round 2: (a) the GEMMs of one eval forward at 4096 x 64 bf16 (multi-wave tiles),
# (b) 3xTF32 GEMM / weight-gradient kernels of the fp32-mode training step.  Each only after the same command has
# exited 0 without ncu.  Reports -> gpurun_out/ncu_r02_*.ncu-rep + raw CSV pages.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export WANDB_MODE=disabled
A="bench.py --mode infer --batch 4096 --frames 64 --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
B="bench.py --precision fp32 --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
timeout 120 python $A > /dev/null 2>&1 && \
timeout 500 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -s 27 -c 9 -o gpurun_out/ncu_r02_infer -f \
  python $A > gpurun_out/ncu_r02_infer.log 2>&1
echo "infer rc=$?"
ncu -i gpurun_out/ncu_r02_infer.ncu-rep --page raw --csv > gpurun_out/ncu_r02_infer_raw.csv 2>/dev/null
timeout 120 python $B > /dev/null 2>&1 && \
timeout 500 ncu --set full --import-source on --clock-control none -k regex:tf32 -s 300 -c 8 -o gpurun_out/ncu_r02_tf32 -f \
  python $B > gpurun_out/ncu_r02_tf32.log 2>&1
echo "tf32 rc=$?"
ncu -i gpurun_out/ncu_r02_tf32.ncu-rep --page raw --csv > gpurun_out/ncu_r02_tf32_raw.csv 2>/dev/null
ls -la gpurun_out | grep ncu_r02
