#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --workload synthM --no-cpu-baseline --no-e2e > gpurun_out/c42_synthM.json 2> gpurun_out/c42_synthM.err
python - <<PY
import json
d = json.load(open("gpurun_out/c42_synthM.json"))
print("synthM", d.get("value"), d.get("ms_per_step"), d["config"].get("launch"))
PY
