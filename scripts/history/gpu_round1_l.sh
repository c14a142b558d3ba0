set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py -q --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_mc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mc.log
tail -8 gpurun_out/pytest_mc.log
B="python bench.py --no-cpu-baseline --no-e2e"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name: $(tail -1 gpurun_out/$name.log | cut -c1-120)"; }
SM="--workload synthM --steps 2 --warmup 1 --chains 50 --nprod 400 --nequil 100"
PCE_MC_TPL=1 run m_k2_t1 $B $SM --kernel 2
PCE_MC_TPL=2 run m_k2_t2 $B $SM --kernel 2
PCE_MC_TPL=1 PCE_DW_BUDGET_MB=0 run m_k2_t1_nodw $B $SM --kernel 2
PCE_MC_TPL=2 run m_k3_t2 $B $SM --kernel 3
PCE_MC_TPL=4 run m_k3_t4 $B $SM --kernel 3
PCE_MC_TPL=2 PCE_DW_BUDGET_MB=0 run m_k3_t2_nodw $B $SM --kernel 3
PCE_MC_TPL=1 run i_k2_t1 $B --workload ifp20 --kernel 2 --steps 2 --warmup 1
PCE_MC_TPL=1 run i_k3_t1 $B --workload ifp20 --kernel 3 --steps 2 --warmup 1
