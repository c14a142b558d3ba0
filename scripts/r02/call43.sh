#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_analytic.py tests/test_gpu_parity_synthetic.py -m gpu -q --timeout 600 -p no:cacheprovider -k "not synth_m and not ifp20" > gpurun_out/c43_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c43_tests.log
tail -3 gpurun_out/c43_tests.log
timeout 300 python bench.py --workload synthA --no-cpu-baseline > gpurun_out/c43_synthA.json 2> gpurun_out/c43_synthA.err
python - <<PY
import json
d = json.load(open("gpurun_out/c43_synthA.json"))
print("synthA", d.get("value"), (d.get("e2e") or {}).get("value"), d.get("ms_per_step"))
PY
