cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD="python bench.py --no-cpu-baseline --no-e2e --workload ifp20 --steps 1 --warmup 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mc_warp -s 1 -c 1 -o gpurun_out/prof_mc_warp_ifp20_v8 $CMD > gpurun_out/ncu_full_w2.log 2>&1
PCETK_B200_LIBRARY=$GRAFT_REPO_ROOT/gpurun_exp/lib_old.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mc_warp -s 1 -c 1 -o gpurun_out/prof_mc_warp_ifp20_v7b $CMD > gpurun_out/ncu_full_w3.log 2>&1
tail -1 gpurun_out/ncu_full_w2.log gpurun_out/ncu_full_w3.log
