cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name: $(grep '^{' gpurun_out/$name.log | cut -c1-90)"; }
run c_new_i1 $B --workload ifp20 --steps 2 --warmup 1
PCETK_B200_LIBRARY=$GRAFT_REPO_ROOT/gpurun_exp/lib_old.so run c_old_i1 $B --workload ifp20 --steps 2 --warmup 1
run c_new_m $B --workload synthM --steps 2 --warmup 1
PCETK_B200_LIBRARY=$GRAFT_REPO_ROOT/gpurun_exp/lib_old.so run c_old_m $B --workload synthM --steps 2 --warmup 1
PCE_MC_WARPS=7 run c_new_m7 $B --workload synthM --steps 2 --warmup 1 --chains 50
run c_new_m4w $B --workload synthM --steps 2 --warmup 1 --waves 4
PCETK_B200_LIBRARY=$GRAFT_REPO_ROOT/gpurun_exp/lib_old.so run c_old_m4w $B --workload synthM --steps 2 --warmup 1 --waves 4
