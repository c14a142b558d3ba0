# round 2, call 4: new enumeration kernel -- parity, synth-A bench, launch list, one full ncu capture of k_enum
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_analytic.py -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r02_pytest4a.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest4a.log
tail -15 gpurun_out/r02_pytest4a.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_pytest4.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest4.log
tail -5 gpurun_out/r02_pytest4.log
timeout 300 python bench.py --workload synthA --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_synthA_a.log 2>&1; cut -c1-400 gpurun_out/r02_bench_synthA_a.log | tail -2
CMD="python bench.py --workload synthA --no-cpu-baseline --no-e2e --steps 1 --warmup 1"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_synthA_v1.csv $CMD > gpurun_out/r02_ncu_list_synthA.log 2>&1; tail -1 gpurun_out/r02_ncu_list_synthA.log | cut -c1-100
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_enum -s 2 -c 2 -o gpurun_out/r02_prof_k_enum_v1 $CMD > gpurun_out/r02_ncu_full_synthA.log 2>&1; tail -1 gpurun_out/r02_ncu_full_synthA.log
