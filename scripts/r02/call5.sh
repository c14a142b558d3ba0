# round 2, call 5: enumeration kernel v2 (two CTAs per SM, staged static program, segmented reduce)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_analytic.py -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r02_pytest5a.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest5a.log
tail -4 gpurun_out/r02_pytest5a.log
timeout 300 python bench.py --workload synthA --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_synthA_b.log 2>&1; cut -c1-330 gpurun_out/r02_bench_synthA_b.log | tail -2
grep -o '"e2e": {[^}]*}' gpurun_out/r02_bench_synthA_b.log
CMD="python bench.py --workload synthA --no-cpu-baseline --no-e2e --steps 1 --warmup 1"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_synthA_v2.csv $CMD > gpurun_out/r02_ncu_list_synthA.log 2>&1; tail -1 gpurun_out/r02_ncu_list_synthA.log | cut -c1-100
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'^k_enum$' -s 2 -c 1 -o gpurun_out/r02_prof_k_enum_v2 $CMD > gpurun_out/r02_ncu_full_synthA.log 2>&1; tail -1 gpurun_out/r02_ncu_full_synthA.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_enum_static' -c 1 -o gpurun_out/r02_prof_k_enum_static_v2 $CMD > gpurun_out/r02_ncu_full_static.log 2>&1; tail -1 gpurun_out/r02_ncu_full_static.log
