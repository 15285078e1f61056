set -e
cd $GRAFT_REPO_ROOT
QUIET=1 timeout 200 python tools/microbench.py G_eval.gemm.conv5 G_train.gemm.decoder.9 G_train.wgrad.conv5 G_train.dgrad.conv6 G_train.wgrad.decoder.9 > /dev/null 2>&1
timeout 500 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"gemm_tc|wgrad_tc" -s 10 -c 12 -o /tmp/g_rep -f python tools/microbench.py G_eval.gemm.conv5 G_train.gemm.decoder.9 G_train.wgrad.conv5 G_train.dgrad.conv6 G_train.wgrad.decoder.9 > gpurun_out/ncu_gemm.log 2>&1
ncu -i /tmp/g_rep.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -3
