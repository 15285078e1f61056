# CTA pairs (cta_group::2) for the 256-column bf16 tile: replay parity at the benchmarked shapes and at small forced
# shapes, then the bench configurations one by one with and without pairs (each under its own timeout).
# usage (under gpurun): bash tools/pair_call.sh TAG
TAG=${1:-pair}
mkdir -p gpurun_out
export WANDB_MODE=disabled
L=gpurun_out/pair_$TAG.log
: > $L
run() { echo "== $*" | tee -a $L; timeout 200 "$@" >> $L 2>&1; echo "rc=$?" | tee -a $L; }
export B2H_PAIR=1
run python -m pytest tests/test_gpu_replay.py -x -q -p no:cacheprovider -k "test_generator_eval_replay_benched_shapes and 256-64-1"
run python -m pytest tests/test_gpu_replay.py -q -p no:cacheprovider -k "256-64-1 or 64-1024-1 or test_benched_step or test_forced_tile_widths_replay"
run python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -q -p no:cacheprovider -k "bf16 or pipelined"
grep -E "passed|failed" $L | tail -8
COMMON="--steps 20 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-extra-configs"
for cfg in "" "--mode infer --batch 4096 --frames 64" "--mode infer --batch 64 --frames 1024" "--feats" "--variant b2h --feats" \
           "--mode infer --variant v2 --feats --batch 4096 --frames 64"; do
  for pair in 1 0; do
    B2H_PAIR=$pair timeout 120 python bench.py $cfg $COMMON > gpurun_out/pair_last.out 2>/dev/null
    echo "pair=$pair [$cfg] rc=$? $(python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/pair_last.out') if l.startswith('{')][0]); print('ms', round(d['ms_per_step'],4), 'value', round(d['value']))
except Exception as e: print('none')
")" | tee -a $L
  done
done
