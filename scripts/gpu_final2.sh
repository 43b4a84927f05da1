set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gfi_gpu.py -x -q 2>&1 | tail -12
