set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_analytic.py -q --timeout 400 -p no:cacheprovider -x > gpurun_out/pytest4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest4.log
tail -15 gpurun_out/pytest4.log
B="python bench.py --no-cpu-baseline"
$B --no-e2e --workload synthA --steps 3 --warmup 1 > gpurun_out/bench4_synthA.log 2>&1; tail -1 gpurun_out/bench4_synthA.log | cut -c1-300
CMD2="python bench.py --no-cpu-baseline --no-e2e --workload synthA --steps 1 --warmup 1"
$CMD2 > gpurun_out/plain_ncu2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_synthA_v4.csv $CMD2 > gpurun_out/ncu_list4.log 2>&1
$CMD2 > gpurun_out/plain_ncu2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_enum$ -s 2 -c 1 -o gpurun_out/prof_enum_v4 $CMD2 > gpurun_out/ncu_full4b.log 2>&1
