# round 2, call 11: chain-per-thread kernel with 247 B per lysozyme chain (28 warps per SM?) -- parity and throughput
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r02_pytest11.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest11.log
tail -8 gpurun_out/r02_pytest11.log
b() { label=$1; shift; env "$@" > gpurun_out/r02_thread_$label.log 2>&1; python - "$label" <<'PY'
import json, sys
d = json.loads([l for l in open("gpurun_out/r02_thread_%s.log" % sys.argv[1]) if l.startswith("{")][-1])
print("%-16s %.4e moves/s  %.1f ms  launch %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["config"]["launch"]))
PY
}
b lyso_auto timeout 300 python bench.py --workload lysozyme --no-e2e --no-cpu-baseline --steps 3 --warmup 2
b lyso_768 PCE_MC_BLOCK=768 timeout 300 python bench.py --workload lysozyme --no-e2e --no-cpu-baseline --steps 3 --warmup 2
b lyso_832 PCE_MC_BLOCK=832 timeout 300 python bench.py --workload lysozyme --no-e2e --no-cpu-baseline --steps 3 --warmup 2
b defensin_auto timeout 300 python bench.py --workload defensin --no-e2e --no-cpu-baseline --steps 3 --warmup 2
