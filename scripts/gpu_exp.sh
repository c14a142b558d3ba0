cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mc.py tests/test_gpu_edge.py -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -1
