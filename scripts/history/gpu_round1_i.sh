set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py -q --timeout 400 -p no:cacheprovider -x > gpurun_out/pytest_mc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mc.log
tail -8 gpurun_out/pytest_mc.log
B="python bench.py --no-cpu-baseline --no-e2e"
for T in 1 2 4; do
PCE_MC_TPL=$T timeout 600 $B --workload synthM --steps 2 --warmup 1 --chains 25 --nprod 100 --nequil 100 > gpurun_out/bench_synthM_tpl$T.log 2>&1; tail -1 gpurun_out/bench_synthM_tpl$T.log | cut -c1-250
PCE_MC_TPL=$T timeout 600 $B --workload ifp20 --kernel 2 --steps 2 --warmup 1 > gpurun_out/bench_ifp20_k2_tpl$T.log 2>&1; tail -1 gpurun_out/bench_ifp20_k2_tpl$T.log | cut -c1-250
done
PCE_MC_TPL=2 timeout 600 $B --workload lysozyme --kernel 2 --steps 2 --warmup 1 --nprod 2000 > gpurun_out/bench_lyso_k2.log 2>&1; tail -1 gpurun_out/bench_lyso_k2.log | cut -c1-250
timeout 600 $B --workload synthM --steps 2 --warmup 1 --chains 50 --nprod 1000 --nequil 100 > gpurun_out/bench_synthM_long2.log 2>&1; tail -1 gpurun_out/bench_synthM_long2.log | cut -c1-250
grep -o '"acceptance": {[^}]*}' gpurun_out/bench_*tpl2.log gpurun_out/bench_lyso_k2.log
