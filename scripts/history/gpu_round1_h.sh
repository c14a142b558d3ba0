set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py tests/test_gpu_analytic.py -q --timeout 400 -p no:cacheprovider -x > gpurun_out/pytest_mc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mc.log
tail -8 gpurun_out/pytest_mc.log
B="python bench.py --no-cpu-baseline --no-e2e"
$B --workload synthM --steps 2 --warmup 1 --chains 25 --nprod 100 --nequil 100 > gpurun_out/bench_synthM_long.log 2>&1; tail -1 gpurun_out/bench_synthM_long.log | cut -c1-200
$B --workload synthA --steps 3 --warmup 1 > gpurun_out/bench_synthA.log 2>&1; tail -1 gpurun_out/bench_synthA.log | cut -c1-200
$B --workload ifp20 --kernel 2 --steps 2 --warmup 1 > gpurun_out/bench_ifp20_k2.log 2>&1; tail -1 gpurun_out/bench_ifp20_k2.log | cut -c1-200
CMD="python bench.py --no-cpu-baseline --no-e2e --workload synthM --steps 1 --warmup 1 --chains 25 --nprod 20 --nequil 20"
$CMD > gpurun_out/plain_ncu.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_mc_warp -s 1 -c 1 -o gpurun_out/prof_mc_warp_v4 $CMD > gpurun_out/ncu_full_w.log 2>&1
