#!/bin/bash
# warp kernel: trials decoded at refill time: MC parity tests, then the default titration curve, ifp20 and synth-M
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py tests/test_gpu_parity_synthetic.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c26_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c26_tests.log
tail -4 gpurun_out/c26_tests.log
timeout 600 python bench.py --workload lysozyme --no-cpu-baseline --no-side-workloads > gpurun_out/c26_lysozyme.json 2> gpurun_out/c26_lysozyme.err
python - <<PY
import json
d = json.load(open("gpurun_out/c26_lysozyme.json"))
print("lysozyme", d.get("value"), d.get("e2e"), d.get("titration_curve"))
PY
for w in ifp20 synthM; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/c26_$w.json 2> gpurun_out/c26_$w.err
  python - <<PY
import json
d = json.load(open("gpurun_out/c26_$w.json"))
print("$w", d.get("value"), d.get("ms_per_step"), d["config"].get("launch"))
PY
done
