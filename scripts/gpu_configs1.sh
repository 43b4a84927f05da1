set -x
cd $GRAFT_REPO_ROOT
timeout 300 python scripts/bench_configs.py mh 2>&1 | tail -2
timeout 300 python scripts/bench_configs.py hmc 2>&1 | tail -2
timeout 300 python scripts/bench_configs.py hmm 2>&1 | tail -2
