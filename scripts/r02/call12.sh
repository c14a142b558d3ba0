cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python scripts/r02/debug_synthm.py 1000 100 400 16; PCE_DW_BUDGET_MB=0 PCE_CROSS_TABLE=0 timeout 600 python scripts/r02/debug_synthm.py 1000 100 400 16) > gpurun_out/r02_debug_synthm.txt 2>&1
cat gpurun_out/r02_debug_synthm.txt
