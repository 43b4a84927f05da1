cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_zz_mv_normal_gpu.py -x -q 2>&1 | tail -25
