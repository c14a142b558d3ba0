cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name: $(grep '^{' gpurun_out/$name.log | cut -c1-90)"; }
run p_i $B --workload ifp20 --steps 2 --warmup 1
run p_m $B --workload synthM --steps 2 --warmup 1
