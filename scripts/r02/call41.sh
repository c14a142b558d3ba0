#!/bin/bash
# shipped tree: whole GPU suite, v8 digests of the small-field builds of the chain-per-warp kernel, default bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/c41_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c41_tests.log
tail -4 gpurun_out/c41_tests.log
TAG=v8
run() { name=$1; kernel=$2; shift 2
  timeout 300 python bench.py "$@" --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/r02_prof_${name}_${TAG}.json 2> gpurun_out/r02_prof_${name}_${TAG}.err || { echo "$name failed"; tail -3 gpurun_out/r02_prof_${name}_${TAG}.err; return; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^${kernel}\$" -s 1 -c 1 -o gpurun_out/r02_prof_${name}_${TAG} python bench.py "$@" --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/r02_ncu_${name}_${TAG}.log 2>&1
  tail -1 gpurun_out/r02_ncu_${name}_${TAG}.log | cut -c1-120
  python scripts/ncu_summary.py gpurun_out/r02_prof_${name}_${TAG}.ncu-rep gpurun_out/r02_${name}_${TAG}.txt > /dev/null 2>&1
  python scripts/ncu_source_lines.py gpurun_out/r02_prof_${name}_${TAG}.ncu-rep 80 > gpurun_out/r02_${name}_${TAG}.lines.txt 2>&1
  rm -f gpurun_out/r02_prof_${name}_${TAG}.ncu-rep
}
run k_mc_warp_ifp20 k_mc_warp --workload ifp20 --nprod 400 --nequil 100 --chains 86
run k_mc_warp_lysozyme64 k_mc_warp --workload lysozyme --nprod 300 --nequil 50 --chains 64
grep -h "issue_active\|inst_executed.sum\|gpu__time\|no_instruction" gpurun_out/r02_k_mc_warp_*_v8.txt
timeout 900 python bench.py > gpurun_out/c41_bench.json 2> gpurun_out/c41_bench.err; tail -c 200 gpurun_out/c41_bench.json
