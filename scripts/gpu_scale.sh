set -x
cd $GRAFT_REPO_ROOT
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g${N}_global.json 2> gpurun_out/bench_g${N}_global.err; tail -3 gpurun_out/bench_g${N}_global.err; cat gpurun_out/bench_g${N}_global.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --multi-gpu independent > gpurun_out/bench_g${N}_indep.json 2> gpurun_out/bench_g${N}_indep.err; tail -3 gpurun_out/bench_g${N}_indep.err; cat gpurun_out/bench_g${N}_indep.json

