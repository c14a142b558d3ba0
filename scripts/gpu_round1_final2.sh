# final evidence run of round 1 after k_mc_thread v10: full GPU test suite, smoke, both bench arms, defensin, ncu capture + launch list
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -3 gpurun_out/pytest_all.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run() { name=$1; shift; local t0=$SECONDS; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name ($((SECONDS-t0)) s): $(grep '^{' gpurun_out/$name.log | cut -c1-100)"; }
run bench_ref python bench.py --impl reference --steps 1 --warmup 1
run bench_default python bench.py
run bench_defensin_full python bench.py --workload defensin --steps 3 --warmup 3 --no-cpu-baseline
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --nprod 200 --nequil 50 --waves 1"
timeout 100 ncu --set full --clock-control none --import-source on -k regex:k_mc_thread -s 1 -c 1 -o gpurun_out/prof_mc_thread_v10 $CMD > gpurun_out/ncu_full_v10.log 2>&1; tail -1 gpurun_out/ncu_full_v10.log
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_lysozyme_v10.csv python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > gpurun_out/ncu_list_lysozyme.log 2>&1; tail -1 gpurun_out/ncu_list_lysozyme.log | cut -c1-80
