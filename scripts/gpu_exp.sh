cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name: $(grep '^{' gpurun_out/$name.log | cut -c1-90)"; }
for w in 7 6 5 4 2; do PCE_MC_WARPS=$w run w_$w $B --workload synthM --steps 2 --warmup 1 --chains $((w*7)) --nprod 400 --nequil 100; done
