set -e
cd $GRAFT_REPO_ROOT
QUIET=1 timeout 200 python tools/microbench.py bn_bwd.conv6 apply.conv6 > /dev/null 2>&1
timeout 500 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"bn_bwd|bn_apply" -s 20 -c 6 -o /tmp/bn_rep -f python tools/microbench.py bn_bwd.conv6 apply.conv6 > gpurun_out/ncu_bn.log 2>&1
ncu -i /tmp/bn_rep.ncu-rep --page raw --csv > gpurun_out/ncu_bn_raw.csv 2>/dev/null
ncu -i /tmp/bn_rep.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_bn_source.csv 2>/dev/null || true
ls -la gpurun_out/ | tail -5
