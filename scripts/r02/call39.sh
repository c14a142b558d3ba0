#!/bin/bash
# small-field warp build: field vector padded to whole warps (ifp20 192 instead of 256 slots per accepted move)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c39_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c39_tests.log
tail -3 gpurun_out/c39_tests.log
timeout 600 python bench.py --workload ifp20 --no-cpu-baseline > gpurun_out/c39_ifp20.json 2> gpurun_out/c39_ifp20.err
python - <<PY
import json
d = json.load(open("gpurun_out/c39_ifp20.json"))
print("ifp20", d.get("value"), (d.get("e2e") or {}).get("value"), d.get("ms_per_step"), d["config"].get("launch"))
PY
timeout 600 python bench.py --workload lysozyme --no-cpu-baseline --no-side-workloads > gpurun_out/c39_lysozyme.json 2> gpurun_out/c39_lysozyme.err
python - <<PY
import json
d = json.load(open("gpurun_out/c39_lysozyme.json"))
print("lysozyme", d.get("value"), d["titration_curve"]["defaults"]["wall_ms_runs"], d["titration_curve"]["reference_effort"]["wall_ms"])
PY
