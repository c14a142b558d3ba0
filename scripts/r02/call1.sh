cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_call1_smi.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -x > gpurun_out/r02_pytest1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest1.log
tail -3 gpurun_out/r02_pytest1.log
timeout 900 python scripts/r02/planner_sweep.py lysozyme 2000 > gpurun_out/r02_planner_lysozyme.txt 2>&1
timeout 400 env SWEEP_CHAINS=1,16,64,256,1024 python scripts/r02/planner_sweep.py defensin 2000 > gpurun_out/r02_planner_defensin.txt 2>&1
tail -5 gpurun_out/r02_planner_lysozyme.txt
