set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -x > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -8 gpurun_out/pytest_all.log
B="python bench.py --no-cpu-baseline --no-e2e"
$B --steps 3 --warmup 2 > gpurun_out/bench_lyso.log 2>&1; tail -1 gpurun_out/bench_lyso.log | cut -c1-200
$B --workload ifp20 --steps 2 --warmup 1 > gpurun_out/bench_ifp20.log 2>&1; tail -1 gpurun_out/bench_ifp20.log | cut -c1-200
$B --workload synthM --steps 2 --warmup 1 --chains 25 --nprod 20 --nequil 20 > gpurun_out/bench_synthM.log 2>&1; tail -1 gpurun_out/bench_synthM.log | cut -c1-200
$B --workload synthM --steps 2 --warmup 1 --chains 25 --nprod 100 --nequil 100 > gpurun_out/bench_synthM_long.log 2>&1; tail -1 gpurun_out/bench_synthM_long.log | cut -c1-200
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --nprod 200 --nequil 50"
$CMD > gpurun_out/plain_ncu.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_mc_thread -s 1 -c 1 -o gpurun_out/prof_mc_thread_v4 $CMD > gpurun_out/ncu_full.log 2>&1
CMD="python bench.py --no-cpu-baseline --no-e2e --workload synthM --steps 1 --warmup 1 --chains 25 --nprod 20 --nequil 20"
$CMD > gpurun_out/plain_ncu.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_mc_warp -s 1 -c 1 -o gpurun_out/prof_mc_warp_v2 $CMD > gpurun_out/ncu_full_w.log 2>&1
