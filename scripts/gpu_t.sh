cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_golden_fixtures.py -x -q -m gpu 2>&1 | tail -25
