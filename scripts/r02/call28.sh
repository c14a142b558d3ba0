#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c28_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c28_tests.log
tail -3 gpurun_out/c28_tests.log
timeout 600 python bench.py --workload lysozyme --no-cpu-baseline --no-side-workloads > gpurun_out/c28_lysozyme.json 2> gpurun_out/c28_lysozyme.err
python - <<PY
import json
d = json.load(open("gpurun_out/c28_lysozyme.json"))
print("lysozyme", d.get("value"), d["titration_curve"]["defaults"]["wall_ms"], d["titration_curve"]["reference_effort"]["wall_ms"])
PY
