# 2 GPUs: NCCL protocol / algorithm for the two small all-reduces of the step (8.96 MB + 0.49 MB fp32), and the fused
# exchange, each as a plain bench run (no extras)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export WANDB_MODE=disabled
L=gpurun_out/nccl_n2_sweep.log
: > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
COMMON="--gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
port=29600
for env in "" "NCCL_PROTO=Simple" "NCCL_PROTO=LL128" "NCCL_ALGO=Tree" "NCCL_NVLS_ENABLE=0" "NCCL_MAX_NCHANNELS=4" "NCCL_MIN_NCHANNELS=16" "B2H_FUSED_DP=1"; do
  port=$((port+1))
  env $env timeout 150 $TR --master-port $port bench.py $COMMON > gpurun_out/n2_last.out 2>/dev/null
  echo "[$env] rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/n2_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'sync', d['ranks_in_sync'])
except Exception as e: print('none')
")" | tee -a $L
done
