cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mc.py -q --timeout 300 -p no:cacheprovider -x -k "specification" 2>&1 | tail -1 > gpurun_out/spec.log; cat gpurun_out/spec.log
M="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active"
for b in 768 832; do
PCE_MC_BLOCK=$b timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > gpurun_out/t_l_$b.log 2>&1; echo "block $b: $(grep '^{' gpurun_out/t_l_$b.log | cut -c1-90) $(tail -1 gpurun_out/t_l_$b.log | grep -o 'Error.*' | cut -c1-100)"
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --nprod 200 --nequil 50 --waves 1"
PCE_MC_BLOCK=$b timeout 600 ncu --metrics $M --clock-control none -k regex:k_mc_thread -s 1 -c 1 --csv --log-file gpurun_out/blk_$b.csv $CMD > /dev/null 2>&1
done
