cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py -q --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_mc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mc.log
tail -4 gpurun_out/pytest_mc.log
B="python bench.py --no-cpu-baseline --no-e2e"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name: $(grep '^{' gpurun_out/$name.log | cut -c1-90)"; }
SM="--workload synthM --steps 2 --warmup 1 --chains 50 --nprod 400 --nequil 100 --kernel 2"
PCE_MC_TPL=1 run y_m_t1 $B $SM
PCE_MC_TPL=2 run y_m_t2 $B $SM
PCE_MC_TPL=4 run y_m_t4 $B $SM
PCE_MC_TPL=1 run y_i_t1 $B --workload ifp20 --kernel 2 --steps 2 --warmup 1
PCE_MC_TPL=2 run y_i_t2 $B --workload ifp20 --kernel 2 --steps 2 --warmup 1
