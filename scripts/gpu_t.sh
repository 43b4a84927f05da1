cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_pf_gpu.py -x -q -k "fused_mass" 2>&1 | tail -25
