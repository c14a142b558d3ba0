#!/bin/bash
# thread kernel front-end rewrite (move table, row table, explicit shared addressing): whole GPU suite, then benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c21_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c21_tests.log
tail -6 gpurun_out/c21_tests.log
for w in lysozyme defensin ifp20; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/c21_$w.json 2> gpurun_out/c21_$w.err
  python - <<PY
import json
d = json.load(open("gpurun_out/c21_$w.json"))
print("$w", d.get("value"), d.get("ms_per_step"), d["config"].get("launch"))
PY
done
