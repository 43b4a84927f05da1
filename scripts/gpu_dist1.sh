set -x
cd $GRAFT_REPO_ROOT
GJB_TEST_N=20000 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_pf_worker.py 2>&1 | tail -30
