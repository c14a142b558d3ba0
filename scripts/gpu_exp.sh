cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_edge.py -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -5
