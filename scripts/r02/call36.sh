#!/bin/bash
# derived tables keyed on the uploaded content: edge / analytic / MC tests, then the synth-A and lysozyme e2e numbers
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_edge.py tests/test_gpu_analytic.py tests/test_gpu_mc.py tests/test_gpu_examples.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c36_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c36_tests.log
tail -5 gpurun_out/c36_tests.log
for w in synthA lysozyme; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-side-workloads > gpurun_out/c36_$w.json 2> gpurun_out/c36_$w.err
  python - <<PY
import json
d = json.load(open("gpurun_out/c36_$w.json"))
print("$w", d.get("value"), (d.get("e2e") or {}).get("value"), d.get("ms_per_step"))
PY
done
