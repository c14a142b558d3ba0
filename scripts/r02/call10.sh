# round 2, call 10: synth-M experiments (warps per SM, trials per lane), fixed reference-seed tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
b() { label=$1; shift; env "$@" timeout 300 python bench.py --workload synthM --no-e2e --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/r02_synthM_$label.log 2>&1; python - "$label" <<'PY'
import json, sys
d = json.loads([l for l in open("gpurun_out/r02_synthM_%s.log" % sys.argv[1]) if l.startswith("{")][-1])
print("%-12s %.3e moves/s  %.1f ms  launch %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["config"]["launch"]))
PY
}
b warps8 PCE_MC_WARPS=8
b warps9 X=1
b warps9_tpl2 PCE_MC_TPL=2
b warps8_tpl2 PCE_MC_WARPS=8 PCE_MC_TPL=2
b warps6 PCE_MC_WARPS=6
timeout 900 python -m pytest tests/test_gpu_parity_synthetic.py tests/test_gpu_edge.py -q --timeout 600 -p no:cacheprovider -k "200_site or 1000_site or test_gpu_edge" > gpurun_out/r02_pytest10.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest10.log
tail -6 gpurun_out/r02_pytest10.log
