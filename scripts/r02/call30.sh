#!/bin/bash
# experiment: two trials per lane and round (PCE_MC_TPL=2) with the leaner evaluation phase
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tpl in 1 2; do
for w in synthM ifp20; do
  PCE_MC_TPL=$tpl timeout 600 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/c30_${w}_$tpl.json 2> gpurun_out/c30_${w}_$tpl.err
  python - <<PY
import json
d = json.load(open("gpurun_out/c30_${w}_$tpl.json"))
print("tpl $tpl $w", d.get("value"), d.get("ms_per_step"), d["config"].get("launch"))
PY
done
PCE_MC_TPL=$tpl timeout 600 python bench.py --workload lysozyme --no-cpu-baseline --no-side-workloads --no-e2e --steps 1 --warmup 1 > gpurun_out/c30_lyso_$tpl.json 2> gpurun_out/c30_lyso_$tpl.err
python - <<PY
import json
d = json.load(open("gpurun_out/c30_lyso_$tpl.json"))
print("tpl $tpl default curve", d["titration_curve"]["defaults"]["wall_ms"], d["titration_curve"]["reference_effort"]["wall_ms"])
PY
done
