#!/bin/bash
# device-side symmetrisation: full GPU suite + synth-M / lysozyme e2e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/c19_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c19_tests.log
tail -5 gpurun_out/c19_tests.log
timeout 600 python bench.py --workload synthM --no-cpu-baseline > gpurun_out/c19_synthM.json 2> gpurun_out/c19_synthM.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/c19_synthM.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "e2e")})
PY
