# Multi-GPU call for the data-parallel exchange (run under `gpurun --gpus 2 --timeout 900 -- 'bash tools/multi_gpu_first_call.sh'`,
# then again with --gpus 8).  Every step runs under its own timeout: the fused kernel traps after 5-10 s if a peer never
# arrives, and a failing step does not stop the later ones.  Logs: gpurun_out/dp${N}_*.log
cd "${GRAFT_REPO_ROOT:-.}"
N=$(python -c 'import torch; print(torch.cuda.device_count())')
mkdir -p gpurun_out
export WANDB_MODE=disabled
run() {  # name, timeout seconds, command...
  name=$1; shift; t=$1; shift
  echo "=== $name" | tee gpurun_out/dp${N}_$name.log
  timeout "$t" "$@" >> gpurun_out/dp${N}_$name.log 2>&1
  echo "rc=$? ($name)" | tee -a gpurun_out/dp${N}_$name.log
  tail -${TAILN:-6} gpurun_out/dp${N}_$name.log
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
BENCH="bench.py --gpus $N --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
if [ -z "$SKIP_FUSED_CHECKS" ]; then
# 1. the kernel alone, one rank per device inside one process (no torch.distributed, no symmetric memory)
B2H_TEST_MULTI_GPU=1 run two_device_test 120 python -m pytest tests/test_fused_dp.py -x -q -k two_devices
# 2. kernel-level timing: symmetric memory (peer loads / stores and, if available, multimem), then CUDA IPC mapping
run kernel_bench 180 $TR --master-port 29512 tools/dp_adam_bench.py
B2H_DP_PEER=ipc run kernel_bench_ipc 180 $TR --master-port 29513 tools/dp_adam_bench.py
# 3. whole step: parity with the NCCL path and step time — without multicast first, then with
B2H_DP_NO_MULTICAST=1 run step_check_p2p 240 $TR --master-port 29514 tools/dp_fused_check.py
run step_check_mc 240 $TR --master-port 29515 tools/dp_fused_check.py
fi
# 4. the bench line both ways, and the per-kernel share of the step on rank 0 (CUPTI)
run bench_nccl 300 $TR --master-port 29516 $BENCH
B2H_FUSED_DP=1 run bench_fused 300 $TR --master-port 29517 $BENCH
TAILN=30 run profile_nccl 200 $TR --master-port 29518 tools/dp_profile.py
B2H_FUSED_DP=1 TAILN=30 run profile_fused 200 $TR --master-port 29519 tools/dp_profile.py
# 5. BASELINE configs 3 (text) and 4 (image) under data parallelism
run bench_text 300 $TR --master-port 29520 $BENCH --feats
run bench_image 300 $TR --master-port 29521 $BENCH --variant b2h --feats
