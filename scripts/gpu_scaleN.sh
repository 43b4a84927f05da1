# usage: gpurun --gpus N -- 'bash scripts/gpu_scaleN.sh N'   (N = 2, 4, 8)
set -x
cd $GRAFT_REPO_ROOT
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/final_bench_g${N}.json 2> gpurun_out/final_bench_g${N}.err; tail -2 gpurun_out/final_bench_g${N}.err; cat gpurun_out/final_bench_g${N}.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --multi-gpu global > gpurun_out/final_bench_g${N}_global.json 2> gpurun_out/final_bench_g${N}_global.err; tail -2 gpurun_out/final_bench_g${N}_global.err; cat gpurun_out/final_bench_g${N}_global.json
