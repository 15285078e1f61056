# ncu --set full of the dominant full-grid GEMM launch (dgrad of decoder.9) -> gpurun_out/ncu_dominant_raw.csv
set -e
cd $GRAFT_REPO_ROOT
QUIET=1 timeout 200 python tools/microbench.py G_train.dgrad.decoder.9 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k gemm_tc_kernel -s 160 -c 4 -o /tmp/dom_rep -f python tools/microbench.py G_train.dgrad.decoder.9 > gpurun_out/ncu_dominant.log 2>&1
ncu -i /tmp/dom_rep.ncu-rep --page raw --csv > gpurun_out/ncu_dominant_raw.csv 2>/dev/null
ls -la gpurun_out | tail -3
