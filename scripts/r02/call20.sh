#!/bin/bash
# thread kernel with the packed move table: parity first, then the headline and defensin
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c20_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c20_tests.log
tail -5 gpurun_out/c20_tests.log
for w in lysozyme defensin; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/c20_$w.json 2> gpurun_out/c20_$w.err
  python - <<PY
import json
d = json.load(open("gpurun_out/c20_$w.json"))
print("$w", d.get("value"), d.get("ms_per_step"), d["config"].get("launch"), d.get("roofline", {}).get("frac"))
PY
done
