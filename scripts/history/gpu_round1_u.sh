cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -5 gpurun_out/pytest_all.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
B="python bench.py --no-cpu-baseline --no-e2e"
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.log 2>&1; echo "$name: $(grep '^{' gpurun_out/$name.log | cut -c1-90) $(grep -o '"launch": {[^}]*}' gpurun_out/$name.log)"; }
run v_l $B --steps 3 --warmup 2
run v_i $B --workload ifp20 --steps 2 --warmup 1
run v_d $B --workload defensin --steps 2 --warmup 1
run v_m $B --workload synthM --steps 2 --warmup 1
