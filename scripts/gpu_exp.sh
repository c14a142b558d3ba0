cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; shift; local t0=$SECONDS; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name ($((SECONDS-t0)) s): $(grep '^{' gpurun_out/$name.log | cut -c1-100)"; }
run bench_ref python bench.py --impl reference --steps 1 --warmup 1
run bench_default python bench.py
run bench_ifp20_full python bench.py --workload ifp20 --steps 3 --warmup 3
run bench_defensin_full python bench.py --workload defensin --steps 3 --warmup 3
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --nprod 200 --nequil 50 --waves 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mc_thread -s 1 -c 1 -o gpurun_out/prof_mc_thread_v8 $CMD > gpurun_out/ncu_full.log 2>&1
