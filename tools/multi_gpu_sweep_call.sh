# One multi-GPU visit: the exchange sweep (one torchrun job), then the per-kernel share of the step on rank 0 for the
# NCCL exchange and the fused one.  usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/multi_gpu_sweep_call.sh'
cd "${GRAFT_REPO_ROOT:-.}"
N=$(python -c 'import torch; print(torch.cuda.device_count())')
mkdir -p gpurun_out
export WANDB_MODE=disabled
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29530 tools/dp_sweep.py > gpurun_out/dp_sweep_n$N.txt 2> gpurun_out/dp_sweep_n$N.err
echo "sweep rc=$?"; grep '^{' gpurun_out/dp_sweep_n$N.txt
timeout 150 $TR --master-port 29531 tools/dp_profile.py > gpurun_out/dp_profile_nccl_n$N.txt 2>&1
echo "profile nccl rc=$?"; grep -A12 "^world" gpurun_out/dp_profile_nccl_n$N.txt
B2H_FUSED_DP=1 B2H_BUCKETS=${FUSED_BUCKETS:-3} timeout 150 $TR --master-port 29532 tools/dp_profile.py > gpurun_out/dp_profile_fused_n$N.txt 2>&1
echo "profile fused rc=$?"; grep -A12 "^world" gpurun_out/dp_profile_fused_n$N.txt
