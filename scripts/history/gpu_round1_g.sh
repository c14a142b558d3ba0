set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py -q --timeout 400 -p no:cacheprovider -x > gpurun_out/pytest_mc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mc.log
tail -8 gpurun_out/pytest_mc.log
B="python bench.py --no-cpu-baseline --no-e2e"
$B --workload ifp20 --steps 2 --warmup 1 > gpurun_out/bench_ifp20.log 2>&1; tail -1 gpurun_out/bench_ifp20.log | cut -c1-200
$B --workload synthM --steps 2 --warmup 1 --chains 25 --nprod 100 --nequil 100 > gpurun_out/bench_synthM_long.log 2>&1; tail -1 gpurun_out/bench_synthM_long.log | cut -c1-200
python bench.py --steps 3 --warmup 2 > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-2500
