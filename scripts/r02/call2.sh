# round 2, call 2: planner check (auto plan over ensemble sizes), default bench (titration-curve wall times), and one full ncu
# capture of the chain-per-warp kernel on a small lysozyme ensemble (64 chains x 41 pH: the package-default titration curve)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 env SWEEP_CHAINS=1,64,512,1024,2772,5544 python scripts/r02/planner_sweep.py lysozyme 2000 > gpurun_out/r02_planner2_lysozyme.txt 2>&1
grep auto gpurun_out/r02_planner2_lysozyme.txt
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_default_a.log 2>&1; grep -o '"titration_curve": {.*}' gpurun_out/r02_bench_default_a.log | cut -c1-600
grep -o '"value": [0-9.e+]*' gpurun_out/r02_bench_default_a.log | head -2
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --nprod 300 --nequil 50 --chains 64 --kernel 2"
timeout 200 $CMD > gpurun_out/r02_small_warp.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mc_warp -s 1 -c 1 -o gpurun_out/r02_prof_mc_warp_lyso64 $CMD > gpurun_out/r02_ncu_warp_lyso64.log 2>&1; tail -1 gpurun_out/r02_ncu_warp_lyso64.log
