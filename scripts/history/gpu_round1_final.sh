# final evidence run of round 1: full GPU test suite, smoke, the bench lines of every workload, launch lists
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -3 gpurun_out/pytest_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run() { name=$1; shift; local t0=$SECONDS; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name ($((SECONDS-t0)) s): $(grep '^{' gpurun_out/$name.log | cut -c1-100)"; }
run bench_ref python bench.py --impl reference --steps 1 --warmup 1
run bench_default python bench.py
run bench_ifp20_full python bench.py --workload ifp20 --steps 3 --warmup 3
run bench_defensin_full python bench.py --workload defensin --steps 3 --warmup 3
run bench_synthM_full python bench.py --workload synthM --steps 3 --warmup 3
run bench_synthA_full python bench.py --workload synthA --steps 3 --warmup 3
for w in lysozyme synthM ifp20; do
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 1 --workload $w"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${w}_r1.csv $CMD > gpurun_out/ncu_list_$w.log 2>&1
done
