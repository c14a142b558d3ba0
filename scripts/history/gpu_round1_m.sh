cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PCE_MC_BLOCK=512
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --nprod 200 --nequil 50 --waves 1"
timeout 300 $CMD > gpurun_out/plain_ncu.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mc_thread -s 1 -c 1 -o gpurun_out/prof_mc_thread_v5b $CMD > gpurun_out/ncu_full.log 2>&1
grep '^{' gpurun_out/plain_ncu.log | cut -c1-100
