#!/bin/bash
# interleaved copies of the move table (thread kernel); stamps of the small-field warp build back to 32 bits
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c33_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c33_tests.log
tail -3 gpurun_out/c33_tests.log
for w in lysozyme defensin ifp20; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-e2e --no-side-workloads > gpurun_out/c33_$w.json 2> gpurun_out/c33_$w.err
  python - <<PY
import json
d = json.load(open("gpurun_out/c33_$w.json"))
print("$w", d.get("value"), d.get("ms_per_step"), d["config"].get("launch"), (d.get("titration_curve") or {}).get("defaults", {}).get("wall_ms"))
PY
done
