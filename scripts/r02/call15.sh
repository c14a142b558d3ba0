cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_analytic.py tests/test_gpu_mc.py -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_pytest15.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest15.log
tail -4 gpurun_out/r02_pytest15.log
for w in synthA synthA36; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${w}_c.log 2>&1; python - $w <<'PY'
import json, sys
d = json.loads([l for l in open("gpurun_out/r02_bench_%s_c.log" % sys.argv[1]) if l.startswith("{")][-1])
print(sys.argv[1], "%.4e states/s  %.3f ms  e2e %.4e  fp64 frac %.3f" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), d["roofline"]["frac"]))
PY
done
CMD="python bench.py --workload synthA --no-cpu-baseline --no-e2e --steps 1 --warmup 1"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_synthA_v3.csv $CMD > gpurun_out/r02_ncu_list_synthA.log 2>&1; tail -1 gpurun_out/r02_ncu_list_synthA.log | cut -c1-100
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'^k_enum$' -s 2 -c 1 -o gpurun_out/r02_prof_k_enum_v3 $CMD > gpurun_out/r02_ncu_full_synthA.log 2>&1; tail -1 gpurun_out/r02_ncu_full_synthA.log
