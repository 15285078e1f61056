# what the driver runs at round end for N GPUs: the reference arm, then this repo's arm, launched the same way
cd "${GRAFT_REPO_ROOT:-.}"
N=$(python -c 'import torch; print(torch.cuda.device_count())')
mkdir -p gpurun_out
export WANDB_MODE=disabled
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555"; fi
timeout 300 $TR bench.py --impl reference --gpus $N --steps ${REF_STEPS:-4} --warmup 1 > gpurun_out/driver_ref_n$N.json 2> gpurun_out/driver_ref_n$N.err
echo "reference rc=$?"; grep '^{' gpurun_out/driver_ref_n$N.json | cut -c1-400
timeout 600 $TR bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 > gpurun_out/driver_b200_n$N.json 2> gpurun_out/driver_b200_n$N.err
echo "b200 rc=$?"; grep '^{' gpurun_out/driver_b200_n$N.json | cut -c1-3000
tail -5 gpurun_out/driver_b200_n$N.err
