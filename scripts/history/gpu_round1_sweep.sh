set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
# (1) wave / block-size sensitivity (short curves: nprod 2000)
for blk in 256 128 64 32; do
  for ch in 924 1024 1848; do
    echo "== block $blk chains $ch" >> gpurun_out/sweep2.log
    PCE_MC_BLOCK=$blk $B --steps 2 --warmup 1 --nprod 2000 --chains $ch 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])" >> gpurun_out/sweep2.log 2>&1
  done
done
# (2) other workloads
$B --workload ifp20 --steps 2 --warmup 1 > gpurun_out/bench_ifp20.log 2>&1
$B --workload synthM --steps 2 --warmup 1 > gpurun_out/bench_synthM.log 2>&1
$B --workload synthA --steps 2 --warmup 1 > gpurun_out/bench_synthA.log 2>&1
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.log 2>&1
# (3) ncu: launch list then full capture of the MC kernel on a short run
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 0 --nprod 200 --nequil 50 --chains 924"
$CMD > gpurun_out/plain_ncu.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mc_thread -c 1 -o gpurun_out/prof_mc_thread $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
cat gpurun_out/sweep2.log
tail -1 gpurun_out/bench_ifp20.log | cut -c1-400
tail -1 gpurun_out/bench_synthM.log | cut -c1-400
tail -1 gpurun_out/bench_synthA.log | cut -c1-400
tail -1 gpurun_out/bench_ref.log | cut -c1-300
