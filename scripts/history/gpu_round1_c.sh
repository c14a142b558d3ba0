set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -x > gpurun_out/pytest3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest3.log
tail -15 gpurun_out/pytest3.log
B="python bench.py --no-cpu-baseline"
$B --no-e2e --steps 3 --warmup 2 > gpurun_out/bench3_lyso.log 2>&1; tail -1 gpurun_out/bench3_lyso.log | cut -c1-400
$B --no-e2e --workload ifp20 --steps 2 --warmup 1 > gpurun_out/bench3_ifp20.log 2>&1; tail -1 gpurun_out/bench3_ifp20.log | cut -c1-300
$B --no-e2e --workload synthA --steps 3 --warmup 1 > gpurun_out/bench3_synthA.log 2>&1; tail -1 gpurun_out/bench3_synthA.log | cut -c1-300
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --nprod 200 --nequil 50"
$CMD > gpurun_out/plain_ncu.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_mc_thread -s 1 -c 1 -o gpurun_out/prof_mc_thread_v3 $CMD > gpurun_out/ncu_full3.log 2>&1
CMD2="python bench.py --no-cpu-baseline --no-e2e --workload synthA --steps 1 --warmup 1"
$CMD2 > gpurun_out/plain_ncu2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_enum$ -s 2 -c 1 -o gpurun_out/prof_enum_v3 $CMD2 > gpurun_out/ncu_full3b.log 2>&1
$CMD2 > gpurun_out/plain_ncu2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_synthA_v3.csv $CMD2 > gpurun_out/ncu_list3.log 2>&1
tail -2 gpurun_out/ncu_full3.log gpurun_out/ncu_full3b.log
