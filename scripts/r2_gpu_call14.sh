set -x
cd $GRAFT_REPO_ROOT
timeout 300 python scratch/bm_exact_check.py 2>&1 | tail -12
