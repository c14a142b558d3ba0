# round 2, call 3: parity after the planner / cross-scan-round changes, default bench, small-ensemble sweep
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -x > gpurun_out/r02_pytest3.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest3.log
tail -5 gpurun_out/r02_pytest3.log
timeout 600 env SWEEP_CHAINS=1,16,64,128,256,512 python scripts/r02/planner_sweep.py lysozyme 2000 > gpurun_out/r02_planner3_lysozyme.txt 2>&1
grep -E "auto|warp c3" gpurun_out/r02_planner3_lysozyme.txt
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_default_b.log 2>&1; grep -o '"titration_curve": {.*}' gpurun_out/r02_bench_default_b.log | cut -c1-400
grep -o '"value": [0-9.e+]*' gpurun_out/r02_bench_default_b.log | head -2
for w in ifp20 synthM; do timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${w}_b.log 2>&1; echo "$w $(grep -o '"value": [0-9.e+]*' gpurun_out/r02_bench_${w}_b.log | head -2 | tr '\n' ' ')"; done
