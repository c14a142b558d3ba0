set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PCE_MC_TPL=1
CMD="python bench.py --no-cpu-baseline --no-e2e --workload synthM --kernel 2 --steps 1 --warmup 1 --chains 25 --nprod 350 --nequil 50"
timeout 300 $CMD > gpurun_out/plain_ncu.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mc_warp -s 1 -c 1 -o gpurun_out/prof_mc_warp_v6 $CMD > gpurun_out/ncu_full_w.log 2>&1
grep '^{' gpurun_out/plain_ncu.log | cut -c1-150
tail -3 gpurun_out/ncu_full_w.log | cut -c1-200
