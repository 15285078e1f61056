# 4-GPU visit: exchange sweep at 4 ranks and at 2 ranks (first two devices), then the bench line at 4 with its defaults.
# usage: gpurun --gpus 4 --timeout 600 -- 'bash tools/multi_gpu_small_call.sh'
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export WANDB_MODE=disabled CASES=${CASES:-v1:0} EXCHANGES=${EXCHANGES:-nccl:1,fused:1,fused-p2p:1}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 4 --master-port 29540 tools/dp_sweep.py > gpurun_out/dp_sweep_n4.txt 2> gpurun_out/dp_sweep_n4.err
echo "sweep4 rc=$?"; grep '^{' gpurun_out/dp_sweep_n4.txt
CUDA_VISIBLE_DEVICES=0,1 timeout 200 $TR --nproc-per-node 2 --master-port 29541 tools/dp_sweep.py > gpurun_out/dp_sweep_n2.txt 2> gpurun_out/dp_sweep_n2.err
echo "sweep2 rc=$?"; grep '^{' gpurun_out/dp_sweep_n2.txt
timeout 200 $TR --nproc-per-node 4 --master-port 29542 bench.py --gpus 4 --steps 30 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "bench4 rc=$?"; cut -c1-1800 gpurun_out/bench_n4.json
