#!/bin/bash
# thread kernel v3: MC parity tests, benches, then ncu digests of the new kernel (lysozyme, defensin)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py tests/test_gpu_distributed.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c22_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c22_tests.log
tail -4 gpurun_out/c22_tests.log
for w in lysozyme defensin; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/c22_$w.json 2> gpurun_out/c22_$w.err
  python - <<PY
import json
d = json.load(open("gpurun_out/c22_$w.json"))
print("$w", d.get("value"), d.get("ms_per_step"), d["config"].get("launch"))
PY
done
TAG=v3
run() { name=$1; kernel=$2; shift 2
  timeout 300 python bench.py "$@" --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/r02_prof_${name}_${TAG}.json 2> gpurun_out/r02_prof_${name}_${TAG}.err || { echo "$name failed"; tail -3 gpurun_out/r02_prof_${name}_${TAG}.err; return; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^${kernel}\$" -s 1 -c 1 -o gpurun_out/r02_prof_${name}_${TAG} python bench.py "$@" --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/r02_ncu_${name}_${TAG}.log 2>&1
  tail -1 gpurun_out/r02_ncu_${name}_${TAG}.log | cut -c1-120
  python scripts/ncu_summary.py gpurun_out/r02_prof_${name}_${TAG}.ncu-rep gpurun_out/r02_${name}_${TAG}.txt > /dev/null 2>&1
  python scripts/ncu_source_lines.py gpurun_out/r02_prof_${name}_${TAG}.ncu-rep 60 > gpurun_out/r02_${name}_${TAG}.lines.txt 2>&1
  rm -f gpurun_out/r02_prof_${name}_${TAG}.ncu-rep
}
run k_mc_thread_lysozyme k_mc_thread --workload lysozyme --nprod 200 --nequil 50 --waves 1
run k_mc_thread_defensin k_mc_thread --workload defensin --nprod 200 --nequil 50 --waves 1
head -20 gpurun_out/r02_k_mc_thread_lysozyme_v3.txt
