cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -1 > gpurun_out/spec.log; cat gpurun_out/spec.log
B="python bench.py --no-cpu-baseline --no-e2e"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name: $(grep '^{' gpurun_out/$name.log | cut -c1-90)"; }
for c in 2 3 4; do PCE_MC_CTAS=$c run u_i_c$c $B --workload ifp20 --steps 2 --warmup 1; done
run u_m $B --workload synthM --steps 2 --warmup 1
