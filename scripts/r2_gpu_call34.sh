# round 2, GPU call 34 (8 GPUs): phase timeline of the step kernel at 8 ranks (GJB_TRACE build of the bench model): where the step period goes
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GJB_NVCC_EXTRA=-DGJB_TRACE timeout 240 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29571 scratch/trace_step_dist.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\|NCCL version" | tail -150 > gpurun_out/r2c34_trace_8gpu.txt; tail -40 gpurun_out/r2c34_trace_8gpu.txt
