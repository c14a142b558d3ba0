cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -5 gpurun_out/pytest_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
B="python bench.py --no-cpu-baseline --no-e2e"
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.log 2>&1; echo "$name: $(grep '^{' gpurun_out/$name.log | cut -c1-90)"; }
run bench_lyso $B --steps 3 --warmup 2
run bench_ifp20 $B --workload ifp20 --steps 2 --warmup 1
run bench_synthM $B --workload synthM --steps 2 --warmup 1 --chains 50 --nprod 400 --nequil 100
CMD="python bench.py --no-cpu-baseline --no-e2e --workload synthM --steps 1 --warmup 1 --chains 25 --nprod 350 --nequil 50"
timeout 300 $CMD > gpurun_out/plain_ncu.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mc_warp -s 1 -c 1 -o gpurun_out/prof_mc_warp_v7 $CMD > gpurun_out/ncu_full_w.log 2>&1
tail -2 gpurun_out/ncu_full_w.log | cut -c1-100
CMD="python bench.py --no-cpu-baseline --no-e2e --workload ifp20 --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/plain_ncu2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mc_warp -s 1 -c 1 -o gpurun_out/prof_mc_warp_ifp20_v7 $CMD > gpurun_out/ncu_full_w2.log 2>&1
tail -2 gpurun_out/ncu_full_w2.log | cut -c1-100
