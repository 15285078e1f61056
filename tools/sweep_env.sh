# usage: tools/sweep_env.sh VAR v1 v2 ... ; prints ms_per_step of the default bench for each value
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v timeout 200 python bench.py --no-cpu-baseline --no-kernel-breakdown 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$VAR=$v', round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,2), 'Mfps')"
done
