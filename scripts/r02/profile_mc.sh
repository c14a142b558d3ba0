# ncu captures of the Monte Carlo kernels as they ship: k_mc_thread (lysozyme, defensin: one wave), k_mc_warp (synth-M; ifp20;
# lysozyme at the package-default ensemble of 64 chains per pH value).
# For each: the command runs once without ncu (its JSON line gives the trial count of one launch), then under ncu.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-v1}
run() { name=$1; kernel=$2; shift 2
  timeout 300 python bench.py "$@" --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/r02_prof_${name}_${TAG}.json 2> gpurun_out/r02_prof_${name}_${TAG}.err || { echo "$name failed"; tail -3 gpurun_out/r02_prof_${name}_${TAG}.err; return; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^${kernel}\$" -s 1 -c 1 -o gpurun_out/r02_prof_${name}_${TAG} python bench.py "$@" --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/r02_ncu_${name}_${TAG}.log 2>&1
  tail -1 gpurun_out/r02_ncu_${name}_${TAG}.log | cut -c1-120
  # digest on the box (the reports together exceed what gpurun copies back): summary + per-source-line view, then drop the report
  python scripts/ncu_summary.py gpurun_out/r02_prof_${name}_${TAG}.ncu-rep gpurun_out/r02_${name}_${TAG}.txt > /dev/null 2>&1
  python scripts/ncu_source_lines.py gpurun_out/r02_prof_${name}_${TAG}.ncu-rep 60 > gpurun_out/r02_${name}_${TAG}.lines.txt 2>&1
  rm -f gpurun_out/r02_prof_${name}_${TAG}.ncu-rep
}
run k_mc_thread_lysozyme k_mc_thread --workload lysozyme --nprod 200 --nequil 50 --waves 1
run k_mc_thread_defensin k_mc_thread --workload defensin --nprod 200 --nequil 50 --waves 1
run k_mc_warp_synthM k_mc_warp --workload synthM --nprod 100 --nequil 100 --waves 1
run k_mc_warp_ifp20 k_mc_warp --workload ifp20 --nprod 400 --nequil 100 --chains 86
run k_mc_warp_lysozyme64 k_mc_warp --workload lysozyme --nprod 300 --nequil 50 --chains 64
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_lysozyme_${TAG}.csv python bench.py --workload lysozyme --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > gpurun_out/r02_ncu_list_lysozyme.log 2>&1; tail -1 gpurun_out/r02_ncu_list_lysozyme.log | cut -c1-100
