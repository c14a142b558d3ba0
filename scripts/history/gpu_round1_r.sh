cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -3 gpurun_out/pytest_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
run() { name=$1; shift; local t0=$SECONDS; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name ($((SECONDS-t0)) s): $(grep '^{' gpurun_out/$name.log | cut -c1-100)"; }
run bench_ref python bench.py --impl reference --steps 1 --warmup 1
run bench_default python bench.py
run bench_ifp20_full python bench.py --workload ifp20 --steps 3 --warmup 3
run bench_synthM_full python bench.py --workload synthM --steps 3 --warmup 3 --chains 50 --nprod 400 --nequil 100
run bench_synthA_full python bench.py --workload synthA --steps 3 --warmup 3
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_lysozyme_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 1 --workload synthM --chains 25 --nprod 100 --nequil 50"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_synthM_r1.csv $CMD > gpurun_out/ncu_list2.log 2>&1
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 1 --workload ifp20"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ifp20_r1.csv $CMD > gpurun_out/ncu_list3.log 2>&1
ls -la gpurun_out/launches_*_r1.csv
