set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -x > gpurun_out/pytest2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest2.log
tail -15 gpurun_out/pytest2.log
B="python bench.py --no-cpu-baseline"
$B --steps 3 --warmup 2 > gpurun_out/bench2_lyso.log 2>&1; tail -1 gpurun_out/bench2_lyso.log | cut -c1-900
for blk in 128 256 384 480; do
  echo "== block $blk" >> gpurun_out/sweep3.log
  PCE_MC_BLOCK=$blk $B --no-e2e --steps 2 --warmup 1 --nprod 2000 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['launch'])" >> gpurun_out/sweep3.log 2>&1
done
cat gpurun_out/sweep3.log
$B --no-e2e --workload ifp20 --steps 2 --warmup 1 > gpurun_out/bench2_ifp20.log 2>&1; tail -1 gpurun_out/bench2_ifp20.log | cut -c1-500
$B --no-e2e --workload synthM --steps 2 --warmup 1 --chains 25 --nprod 20 --nequil 20 > gpurun_out/bench2_synthM.log 2>&1; tail -1 gpurun_out/bench2_synthM.log | cut -c1-500
$B --workload synthA --steps 3 --warmup 1 > gpurun_out/bench2_synthA.log 2>&1; tail -1 gpurun_out/bench2_synthA.log | cut -c1-500
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --nprod 200 --nequil 50"
$CMD > gpurun_out/plain_ncu.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_mc_thread -s 1 -c 1 -o gpurun_out/prof_mc_thread_v2 $CMD > gpurun_out/ncu_full2.log 2>&1
CMD2="python bench.py --no-cpu-baseline --no-e2e --workload synthA --steps 1 --warmup 1"
$CMD2 > gpurun_out/plain_ncu2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_synthA.csv $CMD2 > gpurun_out/ncu_list2.log 2>&1
tail -2 gpurun_out/ncu_full2.log
