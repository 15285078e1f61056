#!/bin/bash
# One GPU-box visit: the -m gpu tests, the bench line, the in-graph op times and the ncu launch list.
# usage (from the repo root, under gpurun): bash tools/gpu_call.sh TAG [pytest-args...]
TAG=${1:-r02}; shift
mkdir -p gpurun_out
export WANDB_MODE=disabled
rm -f gpurun_out/replay_report.txt gpurun_out/parity_noise_report.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider "$@" > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -25 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_$TAG.json
QUIET=1 timeout 300 python tools/microbench.py "" > gpurun_out/microbench_$TAG.txt 2>&1
tail -12 gpurun_out/microbench_$TAG.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu rc=$?"
