cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_pytest14.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest14.log
tail -8 gpurun_out/r02_pytest14.log
b() { label=$1; shift; env "$@" > gpurun_out/r02_thread_$label.log 2>&1; python - "$label" <<'PY'
import json, sys
d = json.loads([l for l in open("gpurun_out/r02_thread_%s.log" % sys.argv[1]) if l.startswith("{")][-1])
print("%-16s %.4e moves/s  e2e %.4e  %.1f ms  launch %s" % (sys.argv[1], d["value"], (d.get("e2e") or {}).get("value", 0), d["ms_per_step"], d["config"]["launch"]))
PY
}
b lyso_auto3 timeout 300 python bench.py --workload lysozyme --no-cpu-baseline --steps 3 --warmup 2
b lyso_768c PCE_MC_BLOCK=768 timeout 300 python bench.py --workload lysozyme --no-e2e --no-cpu-baseline --steps 3 --warmup 2
b defensin_auto3 timeout 300 python bench.py --workload defensin --no-e2e --no-cpu-baseline --steps 3 --warmup 2
