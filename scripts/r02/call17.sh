cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
t0=$SECONDS; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_pytest17.log 2>&1; echo "pytest exit $? in $((SECONDS-t0)) s" >> gpurun_out/r02_pytest17.log
tail -6 gpurun_out/r02_pytest17.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
t0=$SECONDS; timeout 900 python bench.py > gpurun_out/r02_bench_full_b.log 2> gpurun_out/r02_bench_full_b.err; echo "bench exit $? in $((SECONDS-t0)) s"; tail -3 gpurun_out/r02_bench_full_b.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_full_b.log") if l.startswith("{")][-1])
def show(name, w):
    r = w.get("roofline") or {}
    print("%-9s value %.4e e2e %.4e ms/step %.3f roofline %s frac %s cpu %s" % (name, w["value"], (w.get("e2e") or {}).get("value", 0), w["ms_per_step"],
          r.get("bound"), r.get("frac"), (w.get("cpu_baseline") or {}).get("value")))
show("headline", d)
for k, w in d.get("workloads", {}).items(): show(k, w)
print({k: (v["wall_ms"] if isinstance(v, dict) else v) for k, v in d.get("titration_curve", {}).items()}); print(d.get("measured_peaks")); print(d.get("clocks"))
PY
