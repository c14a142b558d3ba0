# round 2, call 6: the restructured bench (all workloads in one line), GPU tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_pytest6.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest6.log
tail -5 gpurun_out/r02_pytest6.log
t0=$SECONDS; timeout 900 python bench.py > gpurun_out/r02_bench_full_a.log 2> gpurun_out/r02_bench_full_a.err; echo "bench exit $? in $((SECONDS-t0)) s"; tail -3 gpurun_out/r02_bench_full_a.err
python - <<'PY'
import json
line = [l for l in open("gpurun_out/r02_bench_full_a.log") if l.startswith("{")][-1]
d = json.loads(line)
def show(name, w):
    r = w.get("roofline") or {}
    print("%-9s %-28s value %.3e e2e %.3e ms/step %.3f roofline %s frac %s cpu %s" % (name, w["metric"], w["value"], (w.get("e2e") or {}).get("value", 0), w["ms_per_step"],
          r.get("bound"), r.get("frac"), (w.get("cpu_baseline") or {}).get("value")))
show("headline", d)
for k, w in d.get("workloads", {}).items(): show(k, w)
print(d.get("titration_curve")); print(d.get("measured_peaks")); print(d.get("clocks"))
PY
