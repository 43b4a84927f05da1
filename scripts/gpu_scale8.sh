set -x
cd $GRAFT_REPO_ROOT
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g${N}_pull.json 2> gpurun_out/bench_g${N}_pull.err; tail -3 gpurun_out/bench_g${N}_pull.err; cat gpurun_out/bench_g${N}_pull.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 scripts/bench_configs.py hmm 2>&1 | grep config > gpurun_out/hmm_g${N}.json; cat gpurun_out/hmm_g${N}.json
