cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py -q --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_mc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mc.log
tail -30 gpurun_out/pytest_mc.log | cut -c1-200
B="python bench.py --no-cpu-baseline --no-e2e"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name: $(grep '^{' gpurun_out/$name.log | cut -c1-90) $(grep -o '"launch": {[^}]*}' gpurun_out/$name.log)"; tail -2 gpurun_out/$name.log | grep -i error | cut -c1-200; }
run t_l $B --steps 2 --warmup 1
run t_i $B --workload ifp20 --steps 2 --warmup 1
run t_m $B --workload synthM --steps 2 --warmup 1
