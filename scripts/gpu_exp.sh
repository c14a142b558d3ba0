cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name: $(grep '^{' gpurun_out/$name.log | cut -c1-90)"; }
SM="--workload synthM --steps 2 --warmup 1 --chains 50 --nprod 400 --nequil 100 --kernel 2"
export PCE_MC_TPL=1
run x_base $B $SM
for v in ldcg u8 u24; do PCETK_B200_LIBRARY=$GRAFT_REPO_ROOT/gpurun_exp/lib_$v.so run x_$v $B $SM; done
