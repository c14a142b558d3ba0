#!/bin/bash
# no per-lane active flag in the warp kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py tests/test_gpu_parity_synthetic.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c37_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c37_tests.log
tail -4 gpurun_out/c37_tests.log
for w in synthM ifp20; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/c37_$w.json 2> gpurun_out/c37_$w.err
  python - <<PY
import json
d = json.load(open("gpurun_out/c37_$w.json"))
print("$w", d.get("value"), (d.get("e2e") or {}).get("value"), d.get("ms_per_step"), d["config"].get("launch"))
PY
done
timeout 600 python bench.py --workload lysozyme --no-cpu-baseline --no-side-workloads > gpurun_out/c37_lysozyme.json 2> gpurun_out/c37_lysozyme.err
python - <<PY
import json
d = json.load(open("gpurun_out/c37_lysozyme.json"))
print("lysozyme", d.get("value"), d["titration_curve"]["defaults"], d["titration_curve"]["reference_effort"]["wall_ms"])
PY
