# First multi-GPU call for the fused data-parallel exchange (run under `gpurun --gpus 2 --timeout 600 -- 'bash tools/multi_gpu_first_call.sh'`,
# then again with --gpus 8).  Every step runs under its own timeout: the kernel traps after 5-10 s if a peer never
# arrives, and a failing step does not stop the later ones.  Logs: gpurun_out/dpfirst_*.log
cd "${GRAFT_REPO_ROOT:-.}"
N=$(python -c 'import torch; print(torch.cuda.device_count())')
mkdir -p gpurun_out
run() {  # name, timeout seconds, command...
  name=$1; shift; t=$1; shift
  echo "=== $name" | tee gpurun_out/dpfirst_$name.log
  timeout "$t" "$@" >> gpurun_out/dpfirst_$name.log 2>&1
  echo "rc=$? ($name)" | tee -a gpurun_out/dpfirst_$name.log
  tail -6 gpurun_out/dpfirst_$name.log
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
# 0. the kept entry points end to end (single process)
B2H_TEST_ENTRY=1 run entry_points 300 python -m pytest tests/test_entry_points_gpu.py -x -q
# 1. the kernel alone, one rank per device inside one process (no torch.distributed, no symmetric memory)
B2H_TEST_MULTI_GPU=1 run two_device_test 120 python -m pytest tests/test_fused_dp.py -x -q -k two_devices
# 2. kernel-level timing: symmetric memory (peer loads / stores and, if available, multimem), then CUDA IPC mapping
run kernel_bench 180 $TR --master-port 29512 tools/dp_adam_bench.py
B2H_DP_PEER=ipc run kernel_bench_ipc 180 $TR --master-port 29513 tools/dp_adam_bench.py
# 3. whole step: parity with the NCCL path and step time — without multicast first, then with
B2H_DP_NO_MULTICAST=1 run step_check_p2p 240 $TR --master-port 29514 tools/dp_fused_check.py
run step_check_mc 240 $TR --master-port 29515 tools/dp_fused_check.py
# 4. the bench line both ways
run bench_nccl 300 $TR --master-port 29516 bench.py --gpus $N --no-cpu-baseline --no-kernel-breakdown
B2H_FUSED_DP=1 run bench_fused 300 $TR --master-port 29517 bench.py --gpus $N --no-cpu-baseline --no-kernel-breakdown
