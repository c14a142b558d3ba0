set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge.py -q --timeout 300 -p no:cacheprovider -x -k "1000_site" > gpurun_out/pytest_mc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mc.log
tail -8 gpurun_out/pytest_mc.log
B="python bench.py --no-cpu-baseline --no-e2e"
for T in 2 1 4; do
PCE_MC_TPL=$T timeout 300 $B --workload synthM --kernel 3 --steps 2 --warmup 1 --chains 25 --nprod 100 --nequil 100 > gpurun_out/bench_synthM_k3_tpl$T.log 2>&1; tail -1 gpurun_out/bench_synthM_k3_tpl$T.log | cut -c1-250
done
timeout 300 $B --workload synthM --kernel 3 --steps 2 --warmup 1 --chains 50 --nprod 1000 --nequil 100 > gpurun_out/bench_synthM_k3_long.log 2>&1; tail -1 gpurun_out/bench_synthM_k3_long.log | cut -c1-250
