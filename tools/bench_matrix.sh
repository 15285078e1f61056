# the other BASELINE configs (C3 text, C4 image, C5 inference sweep) through bench.py; one JSON line each
cd $GRAFT_REPO_ROOT
run() { timeout 300 python bench.py --no-cpu-baseline --no-kernel-breakdown "$@" 2>/dev/null | tail -1; }
{
run --variant v1 --feats
run --variant b2h --feats
run --variant v4 --feats
run --precision fp32
run --mode infer --batch 256 --frames 64
run --mode infer --batch 1024 --frames 64
run --mode infer --batch 4096 --frames 64
run --mode infer --batch 256 --frames 1024
run --mode infer --variant v2 --feats --batch 1024 --frames 64
# incremental-fingers models (launch_exp_incr_fingers.sh: v2 + text, one model per pipeline)
run --mode infer --variant v2 --feats --pipeline arm_wh2finger1 --batch 1024 --frames 64
run --mode infer --variant v2 --feats --pipeline arm_wh2finger6 --batch 1024 --frames 64
run --mode infer --variant v2 --feats --pipeline arm_wh2finger11 --batch 1024 --frames 64
run --variant v2 --feats --pipeline arm_wh2finger6
} > gpurun_out/bench_matrix_r01.jsonl
python - <<'PY'
import json
for ln in open('gpurun_out/bench_matrix_r01.jsonl'):
    try:
        d=json.loads(ln)
        print(d['config']['workload'][:70], '|', d['dtype'], '|', round(d['ms_per_step'],3), 'ms |', round(d['value']/1e6,2), 'Mfps | e2e', round(d['e2e']['value']/1e6,2))
    except Exception as e:
        print('bad line', ln[:100])
PY
