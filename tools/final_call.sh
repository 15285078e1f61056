#!/bin/bash
# Round-end style visit: smoke(), the -m gpu tests, the full bench line, in-graph op times, ncu launch list.
TAG=${1:-r02_final}
mkdir -p gpurun_out
export WANDB_MODE=disabled
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log
bash tools/gpu_call.sh $TAG
