cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -15 gpurun_out/pytest_all.log
timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_default2.log 2>&1; grep -o '"titration_curve": {.*}' gpurun_out/bench_default2.log | cut -c1-700
