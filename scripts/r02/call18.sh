#!/bin/bash
# full GPU suite with the final tree + synth-A bench after the host-side trims
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c18_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c18_tests.log
tail -5 gpurun_out/c18_tests.log
timeout 300 python bench.py --workload synthA --no-cpu-baseline > gpurun_out/c18_synthA.json 2> gpurun_out/c18_synthA.err; tail -c 1500 gpurun_out/c18_synthA.json
