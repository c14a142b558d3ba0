cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py -q --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_mc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mc.log
tail -4 gpurun_out/pytest_mc.log
B="python bench.py --no-cpu-baseline --no-e2e"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name: $(grep '^{' gpurun_out/$name.log | cut -c1-90)"; }
run q_m $B --workload synthM --steps 2 --warmup 1 --chains 50 --nprod 400 --nequil 100
PCE_CROSS_TABLE=0 run q_m_nocross $B --workload synthM --steps 2 --warmup 1 --chains 50 --nprod 400 --nequil 100
run q_i $B --workload ifp20 --steps 2 --warmup 1
grep -o '"on_chip": {[^}]*}' gpurun_out/q_m.log | tr ',' '\n' | grep l2_
