# round 2, call 9: constant-bank micro-benchmark, distributed GPU test, the fixed synthetic parity tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./gpurun_exp/cq > gpurun_out/r02_cq.txt 2>&1; cat gpurun_out/r02_cq.txt
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_parity_synthetic.py -q --timeout 600 -p no:cacheprovider -k "distributed or 200_site or 1000_sites" > gpurun_out/r02_pytest9.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest9.log
tail -15 gpurun_out/r02_pytest9.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x --deselect tests/test_gpu_parity_synthetic.py > gpurun_out/r02_pytest9b.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest9b.log
tail -5 gpurun_out/r02_pytest9b.log
