set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -25 gpurun_out/pytest_all.log
python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > gpurun_out/bench_lyso.log 2>&1; tail -1 gpurun_out/bench_lyso.log | cut -c1-200
