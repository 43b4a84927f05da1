# round 2, GPU call 25 (2 GPUs): phase timeline of the step kernel with the rank-level table and with the per-tile table (GJB_TRACE build)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for L in 1 0; do
  GJB_STEP_LIGHT=$L GJB_NVCC_EXTRA=-DGJB_TRACE timeout 600 $TR --nproc-per-node 2 --master-port 2956$L scratch/trace_step_dist.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\|NCCL version" | tail -32 | tee gpurun_out/r2c25_trace_2gpu_light$L.txt
done
