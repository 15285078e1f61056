# fp32 mode on the tensor cores (3xTF32): per-kernel replay against the restatements, end-to-end parity, sanitizer, bench.
# usage (under gpurun): bash tools/tf32_call.sh TAG
TAG=${1:-tf32}
mkdir -p gpurun_out
export WANDB_MODE=disabled
rm -f gpurun_out/replay_report.txt gpurun_out/parity_noise_report.txt
L=gpurun_out/tf32_$TAG.log
echo "== first kernel (eval forward, fp32, small)" | tee $L
timeout 180 python -m pytest tests/test_gpu_replay.py -x -q -p no:cacheprovider -k "test_generator_eval_replay and 4-64-0" >> $L 2>&1
echo "rc=$?" | tee -a $L
echo "== train replay small fp32" | tee -a $L
timeout 180 python -m pytest tests/test_gpu_replay.py -x -q -p no:cacheprovider -k "test_generator_train_replay and 4-64-0" >> $L 2>&1
echo "rc=$?" | tee -a $L
grep -A14 "gen-train v1 feats=False 36->252 B=4 T=64 dtype=0" gpurun_out/replay_report.txt | tail -15 | tee -a $L
echo "== all replay + parity + shapes + golden" | tee -a $L
timeout 1200 python -m pytest tests/test_gpu_replay.py tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_golden.py \
  tests/test_modelzoo_shim.py -q -p no:cacheprovider -m gpu > gpurun_out/tf32_pytest_$TAG.log 2>&1
echo "rc=$?" | tee -a $L
tail -15 gpurun_out/tf32_pytest_$TAG.log | tee -a $L
echo "== bench fp32" | tee -a $L
timeout 300 python bench.py --precision fp32 --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/bench_fp32_$TAG.json 2> gpurun_out/bench_fp32_$TAG.err
echo "rc=$?" | tee -a $L
cut -c1-300 gpurun_out/bench_fp32_$TAG.json | tee -a $L
B2H_FP32_SIMT=1 timeout 300 python bench.py --precision fp32 --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs --no-kernel-breakdown > gpurun_out/bench_fp32_simt_$TAG.json 2>/dev/null
cut -c1-300 gpurun_out/bench_fp32_simt_$TAG.json | tee -a $L
