set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
run() { name=$1; shift; local t0=$SECONDS; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name: $(grep '^{' gpurun_out/$name.log | cut -c1-120) wall $((SECONDS - t0)) s"; }
SM="--workload synthM --steps 2 --warmup 1 --chains 50 --nprod 400 --nequil 100"
PCE_MC_TPL=1 run m_k2_t1 $B $SM --kernel 2
PCE_MC_TPL=2 run m_k2_t2 $B $SM --kernel 2
PCE_MC_TPL=4 run m_k2_t4 $B $SM --kernel 2
PCE_MC_TPL=1 PCE_DW_BUDGET_MB=0 run m_k2_t1_nodw $B $SM --kernel 2
PCE_MC_TPL=1 run m_k3_t1 $B $SM --kernel 3
PCE_MC_TPL=2 run m_k3_t2 $B $SM --kernel 3
PCE_MC_TPL=4 run m_k3_t4 $B $SM --kernel 3
PCE_MC_TPL=1 run i_k2_t1 $B --workload ifp20 --kernel 2 --steps 2 --warmup 1
PCE_MC_TPL=2 run i_k2_t2 $B --workload ifp20 --kernel 2 --steps 2 --warmup 1
PCE_MC_TPL=1 run i_k3_t1 $B --workload ifp20 --kernel 3 --steps 2 --warmup 1
PCE_MC_TPL=1 run m_anom $B --workload synthM --kernel 2 --steps 1 --warmup 1 --chains 25 --nprod 150 --nequil 50
