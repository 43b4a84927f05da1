set -x
cd $GRAFT_REPO_ROOT
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_g${N}.json 2> gpurun_out/final_bench_g${N}.err; tail -2 gpurun_out/final_bench_g${N}.err; cat gpurun_out/final_bench_g${N}.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --multi-gpu independent > gpurun_out/final_bench_g${N}_independent.json 2> gpurun_out/final_bench_g${N}_independent.err; cat gpurun_out/final_bench_g${N}_independent.json
