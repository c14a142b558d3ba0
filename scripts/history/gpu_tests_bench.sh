set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -25 gpurun_out/pytest_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > gpurun_out/bench_lyso.log 2>&1; tail -1 gpurun_out/bench_lyso.log | cut -c1-200
python bench.py --no-cpu-baseline --no-e2e --workload ifp20 --steps 2 --warmup 1 > gpurun_out/bench_ifp20.log 2>&1; tail -1 gpurun_out/bench_ifp20.log | cut -c1-200
