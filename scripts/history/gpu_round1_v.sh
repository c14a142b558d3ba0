cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --nprod 200 --nequil 50 --waves 1"
timeout 300 $CMD > gpurun_out/plain_ncu.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mc_thread -s 1 -c 1 -o gpurun_out/prof_mc_thread_v6 $CMD > gpurun_out/ncu_full.log 2>&1
grep '^{' gpurun_out/plain_ncu.log | cut -c1-100
CMD="python bench.py --no-cpu-baseline --no-e2e --workload synthM --steps 1 --warmup 1 --chains 28 --nprod 350 --nequil 50"
timeout 300 $CMD > gpurun_out/plain_ncu2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mc_warp -s 1 -c 1 -o gpurun_out/prof_mc_warp_v8 $CMD > gpurun_out/ncu_full_w.log 2>&1
grep '^{' gpurun_out/plain_ncu2.log | cut -c1-100
for w in lysozyme synthM ifp20; do
CMD="python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 1 --workload $w"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${w}_r1.csv $CMD > gpurun_out/ncu_list_$w.log 2>&1
done
ls -la gpurun_out/launches_*_r1.csv
