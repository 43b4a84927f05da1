set -x
cd $GRAFT_REPO_ROOT
GJB_TEST_N=30000 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_pf_worker.py 2>&1 | grep "DIST_PF\|rror"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g2_global.json 2> gpurun_out/bench_g2_global.err; tail -3 gpurun_out/bench_g2_global.err; cat gpurun_out/bench_g2_global.json
