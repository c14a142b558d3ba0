set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --workload ifp20 --steps 2 --warmup 1"
for k in 1 2; do
  echo "== kernel $k" >> gpurun_out/ifp20_sweep.log
  $B --kernel $k 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['launch'])" >> gpurun_out/ifp20_sweep.log 2>&1
done
for blk in 32 64 96 128; do
  echo "== kernel 1 block $blk" >> gpurun_out/ifp20_sweep.log
  PCE_MC_BLOCK=$blk $B --kernel 1 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['launch'])" >> gpurun_out/ifp20_sweep.log 2>&1
done
cat gpurun_out/ifp20_sweep.log
CMD="python bench.py --no-cpu-baseline --no-e2e --workload ifp20 --steps 1 --warmup 1 --nprod 100 --nequil 50"
$CMD > gpurun_out/plain_ncu.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_mc_thread -s 1 -c 1 -o gpurun_out/prof_mc_thread_ifp20 $CMD > gpurun_out/ncu_full.log 2>&1
