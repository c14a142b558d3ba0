cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; shift; local t0=$SECONDS; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name ($((SECONDS-t0)) s): $(grep '^{' gpurun_out/$name.log | cut -c1-100)"; }
run bench_ref python bench.py --impl reference --steps 1 --warmup 1
run bench_default python bench.py
run bench_ifp20_full python bench.py --workload ifp20 --steps 3 --warmup 3
run bench_synthM_full python bench.py --workload synthM --steps 3 --warmup 3
run bench_synthA_full python bench.py --workload synthA --steps 3 --warmup 3
