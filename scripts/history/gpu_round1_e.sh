set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -x > gpurun_out/pytest5.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest5.log
tail -5 gpurun_out/pytest5.log
B="python bench.py --no-cpu-baseline --no-e2e"
$B --workload synthA --steps 3 --warmup 1 > gpurun_out/bench5_synthA.log 2>&1; tail -1 gpurun_out/bench5_synthA.log | cut -c1-200
for c in 25 50 101; do
$B --workload synthM --steps 2 --warmup 1 --chains $c --nprod 20 --nequil 20 > gpurun_out/bench5_synthM_$c.log 2>&1; tail -1 gpurun_out/bench5_synthM_$c.log | cut -c1-200
done
CMD="python bench.py --no-cpu-baseline --no-e2e --workload synthM --steps 1 --warmup 1 --chains 25 --nprod 4 --nequil 4"
$CMD > gpurun_out/plain_ncu.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_mc_warp -s 1 -c 1 -o gpurun_out/prof_mc_warp_v1 $CMD > gpurun_out/ncu_full5.log 2>&1
tail -2 gpurun_out/ncu_full5.log
